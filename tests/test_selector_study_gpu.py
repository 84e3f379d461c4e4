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
        if oracle.have_ref_selector():  # the reference's own adaptive.cpp compiled in place agrees
            assert oracle.ref_adaptive_choice(d.rowptr.cpu().numpy()) == choice, name
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


def test_selector_study_on_suitesparse_shaped_matrices():
    """The matrices of the reference's evaluation (examples/large-data-set-batch.sh:23-52) are not shipped; stand-ins
    with the documented rows / columns / average row length and a structure of the same family are multiplied with the
    default plan (no per-matrix options), checked against the reference's CPU SpMV on sampled rows, and timed next to
    cuSPARSE on the same buffers. The table (reference selector choice, our form, times) goes to
    gpurun_out/selector_study_suitesparse.json (copied to profiles/ and quoted in DESIGN.md). Asserted: parity on every
    shape; the regular families (FEM-like, banded, short rectangular rows, dense blocks) at least on par with cuSPARSE.
    The skewed row-length families (Ga41As41H72, vas_stokes_2M stand-ins: log-normal rows with a heavy tail, a fifth
    of the columns scattered) are the known weak spot of the MIXED / direct kernels (0.7-0.9 x cuSPARSE, DESIGN.md
    §7): reported, and guarded only against getting worse than 0.6 x."""
    import ctypes as C
    import torch
    from oracle import sampled
    from spmv_acc_b200 import _lib
    table = []
    for name in synth.SUITESPARSE_SHAPES:
        d = synth.suitesparse_like_device(name)
        rows_ref, cols_ref, avg_ref, family = synth.SUITESPARSE_SHAPES[name]
        assert (d.rows, d.cols) == (rows_ref, cols_ref) and abs(d.nnz / d.rows - avg_ref) <= 0.2 * avg_ref, name
        choice = oracle.port_adaptive_choice(d.rowptr.cpu().numpy())
        if oracle.have_ref_selector():
            assert oracle.ref_adaptive_choice(d.rowptr.cpu().numpy()) == choice, name
        plan = SpmvPlan(desc_of(d))
        i = plan.info()
        x, y0 = synth.vector_device(d.cols, 2), synth.vector_device(d.rows, 3)
        y = y0.clone()
        plan.execute(0.75, -0.5, x, y)
        torch.cuda.synchronize()
        lens = d.rowptr[1:] - d.rowptr[:-1]
        rows = sampled.sample_rows(d.rows, 5000, seed=3, must_include=torch.topk(lens, 16).indices.cpu().tolist())
        res = sampled.check_sampled_rows(d.rowptr, d.col, d.val, x, y0, y, 0.75, -0.5, rows)
        assert res["ok"], (name, res)

        def timed(fn, reps=30):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            e1.synchronize()
            return e0.elapsed_time(e1) / reps

        ms = timed(lambda: plan.execute(1.0, 1.0, x, y))
        X = _lib.ctx()
        h = C.c_void_p()
        rc = X.spmv_b200_ctx_cusparse_create(C.byref(h), d.rows, d.cols, d.nnz, d.rowptr.data_ptr(), d.col.data_ptr(),
                                             d.val.data_ptr(), x.data_ptr(), y.data_ptr(), 0)
        assert rc == 0
        st = torch.cuda.current_stream().cuda_stream
        ms_cusparse = timed(lambda: X.spmv_b200_ctx_cusparse_spmv(h, 1.0, 1.0, st))
        X.spmv_b200_ctx_cusparse_destroy(h)
        ours = ("direct: one warp per row block, no shared memory" if i.direct else
                "tiled: " + ", ".join(f"{k} {c}" for k, c in zip(("SHORT", "MEDIUM", "MIXED"), i.tiles_per_kind) if c)
                + (", staged x" if i.xstage else ""))
        b_alg = synth.algorithmic_bytes(d.rows, d.cols, d.nnz)
        table.append({"matrix_shape_of": name, "family": family, "rows": d.rows, "cols": d.cols, "nnz": d.nnz,
                      "reference_selector": choice, "ours": ours, "tile_nnz": i.tile_nnz, "split_rows": i.nsplit_rows,
                      "gather_lines_per_gather": round(i.gather_lines / max(i.gather_active, 1), 3),
                      "ms": round(ms, 5), "effective_gbs": round(b_alg / ms / 1e6, 1),
                      "ms_cusparse": round(ms_cusparse, 5), "speedup_vs_cusparse": round(ms_cusparse / ms, 3),
                      "rows_checked": res["rows_checked"], "worst_error_over_bound": res["worst_error_over_bound"]})
        plan.destroy()
        del d, x, y, y0
        torch.cuda.empty_cache()
    out = ROOT / "gpurun_out"
    if out.is_dir():
        (out / "selector_study_suitesparse.json").write_text(json.dumps(table, indent=1) + "\n")
    slow = [(t["matrix_shape_of"], t["speedup_vs_cusparse"]) for t in table
            if t["speedup_vs_cusparse"] < (0.6 if t["family"] == "skewed" else 0.85)]
    assert not slow, f"default plan too far behind cuSPARSE on {slow}"
