"""Runs a handful of device-resident SpMVs of one BASELINE.json shape (target for `ncu --set full`).

    python tools/profile_one.py <c2|c3|c4|c5|c5s> [--tile T] [--vec-div V] [--flags F] [--reps R]
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import torch  # noqa: E402

from spmv_acc_b200 import CsrDesc, SpmvPlan, make_options, synth  # noqa: E402
from sweep import make  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--vec-div", type=int, default=0)
ap.add_argument("--medium-max", type=int, default=0)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--reps", type=int, default=6)
a = ap.parse_args()
csr = make(a.workload)
plan = SpmvPlan(CsrDesc(csr.rows, csr.cols, csr.nnz, csr.rowptr, csr.col, csr.val),
                make_options(a.tile, 0, a.medium_max, a.vec_div, a.flags))
x = synth.vector_device(csr.cols, 2)
y = synth.vector_device(csr.rows, 3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.reps):
    if i == a.reps - 1:
        e0.record()
    plan.execute(1.0, 1.0, x, y)
e1.record()
torch.cuda.synchronize()
i = plan.info()
print(f"{a.workload}: rows={csr.rows} nnz={csr.nnz} kinds={list(i.tiles_per_kind)} split={i.nsplit_rows} "
      f"last launch {e0.elapsed_time(e1):.4f} ms, B_alg={synth.algorithmic_bytes(csr.rows, csr.cols, csr.nnz)}")
