"""Gather bound with SM affinity (development / evidence tool): uniformly random gathers from a table that does not fit
the L2 when every SM reads all of it (80 MB), with the SMs split into two groups that each read only one half of the
table (spmv_b200_ctx_gather_affine). Several guesses of how SM ids map to the two dies, and a control whose groups
are unrelated to the SM. Prints one JSON line per case.

    python tools/gather_affine.py [--n 10000000] [--nnz 320000000]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from spmv_acc_b200 import _lib, synth  # noqa: E402

MAPS = {0: "smid & 1", 1: "(smid >> 1) & 1", 2: "smid >= #SMs / 2", 3: "((smid >> 1) % 8) < 4", 4: "(smid / 18) & 1",
        9: "control: blockIdx.x & 1"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--nnz", type=int, default=320_000_000)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    X = _lib.ctx()
    half = a.nnz // 2
    x = synth.vector_device(a.n, 2)
    col_lo = torch.randint(0, a.n // 2, (half,), dtype=torch.int32, device="cuda")
    col_hi = torch.randint(a.n // 2, a.n, (half,), dtype=torch.int32, device="cuda")
    want = float(x[col_lo.long()].sum().item() + x[col_hi.long()].sum().item())
    tickets = torch.zeros(2, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for cps in (8, 4):
        grid = X.spmv_b200_ctx_gather_affine(half, 0, 0, 0, 0, 0, 0, cps, 0)
        out = torch.zeros(grid * 256, dtype=torch.float64, device="cuda")
        for m, name in MAPS.items():
            args = (half, col_lo.data_ptr(), col_hi.data_ptr(), x.data_ptr(), out.data_ptr(), tickets.data_ptr(), m, cps, st)
            rc = X.spmv_b200_ctx_gather_affine(*args)
            assert rc == 0, rc
            torch.cuda.synchronize()
            total = float(out.sum().item())
            for _ in range(2):
                X.spmv_b200_ctx_gather_affine(*args)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                X.spmv_b200_ctx_gather_affine(*args)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            print(json.dumps({"x_MB": a.n * 8 / 1e6, "gathers": 2 * half, "groups_by": name, "ctas_per_sm": cps,
                              "ms": round(ms, 4), "Ggather_s": round(2 * half / ms / 1e6, 1),
                              "sum_ok": abs(total - want) <= 1e-9 * max(1.0, abs(want))}), flush=True)
    # the same gathers with no grouping at all (every SM reads the whole table): the round-1 bound kernel
    col = torch.cat([col_lo, col_hi])[torch.randperm(2 * half, device="cuda")]
    grid = X.spmv_b200_ctx_gather_bound(2 * half, 0, 0, 0, 0, 0, 8, 0, 0)
    out = torch.zeros(grid * 256, dtype=torch.float64, device="cuda")
    args = (2 * half, col.data_ptr(), 0, x.data_ptr(), out.data_ptr(), 0, 8, 0, st)
    for _ in range(3):
        X.spmv_b200_ctx_gather_bound(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        X.spmv_b200_ctx_gather_bound(*args)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print(json.dumps({"x_MB": a.n * 8 / 1e6, "gathers": 2 * half, "groups_by": "none (every SM gathers from the whole table)",
                      "ms": round(ms, 4), "Ggather_s": round(2 * half / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
