"""Selector study: what the reference's 4-sample run-time selector would run on each
BASELINE.json shape (oracle.port_adaptive_choice, src/acc/hip-adaptive/adaptive.cpp:16-67), next to what our full row
analysis assigns. The table is written to gpurun_out/selector_study.json (and quoted in DESIGN.md §7)."""
import json
from pathlib import Path

import pytest

import oracle
from gpu_helpers import desc_of
from spmv_acc_b200 import SpmvPlan, synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

CASES = [
    ("C2 5-point 4096^2", lambda: synth.stencil2d_device(4096), "adaptive line"),
    ("C3 uniform 1e7 x 1e7, 32/row", lambda: synth.uniform_device(10_000_000, 10_000_000, 32, seed=1), "adaptive flat"),
    ("C4 R-MAT 2^24, 2^28 nnz", lambda: synth.rmat_device(24, 16, seed=1), "adaptive flat"),
    ("C5s 27-point 256^3", lambda: synth.stencil3d_device(256), "adaptive flat"),
]


def test_reference_selector_next_to_our_analysis():
    import torch
    table = []
    for name, make, expected in CASES:
        d = make()
        choice = oracle.port_adaptive_choice(d.rowptr.cpu().numpy())
        assert choice == expected, (name, choice)
        plan = SpmvPlan(desc_of(d))
        i = plan.info()
        ours = ("direct: one warp per row block, no shared memory" if i.direct else
                "tiled: " + ", ".join(f"{k} {c}" for k, c in zip(("SHORT", "MEDIUM", "MIXED"), i.tiles_per_kind) if c))
        table.append({"matrix": name, "rows": i.m, "nnz": i.nnz, "reference_selector": choice, "ours": ours,
                      "tile_nnz": i.tile_nnz, "split_rows": i.nsplit_rows,
                      "rows_per_bin_short_medium_long_verylong": list(i.bin_rows),
                      "gather_lines_per_gather": round(i.gather_lines / max(i.gather_active, 1), 3)})
        plan.destroy()
        del d
        torch.cuda.empty_cache()
    out = ROOT / "gpurun_out"
    if out.is_dir():
        (out / "selector_study.json").write_text(json.dumps(table, indent=1) + "\n")
    # regular stencils: one kernel kind, no split rows; the power-law matrix is the one that takes the direct form
    assert table[0]["ours"].startswith("tiled: SHORT") and table[3]["ours"].startswith("tiled: MEDIUM")
    assert table[2]["ours"].startswith("direct") and table[1]["ours"].startswith("tiled: MEDIUM")
