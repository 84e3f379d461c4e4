"""GPU tuning sweep (development tool): times the device-resident SpMV for the BASELINE.json shapes over plan options
and prints one JSON line per (workload, options). Not part of the product or of the tests.

    python tools/sweep.py [--workloads c2,c3,c4,c5] [--tiles 1024,2048,4096,8192] [--reps 50]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from spmv_acc_b200 import FLAG_NO_TMA, CsrDesc, SpmvPlan, make_options, synth  # noqa: E402


def make(workload):
    if workload == "c2":
        return synth.stencil2d_device(4096)
    if workload == "c3":
        return synth.uniform_device(10_000_000, 10_000_000, 32, seed=1)
    if workload == "c4":
        return synth.rmat_device(24, 16, seed=1)
    if workload == "c5":
        return synth.stencil3d_device(384)
    if workload == "c5s":
        return synth.stencil3d_device(256)
    raise ValueError(workload)


def time_plan(csr, opt, reps, alpha=1.0, beta=1.0):
    plan = SpmvPlan(CsrDesc(csr.rows, csr.cols, csr.nnz, csr.rowptr, csr.col, csr.val), opt)
    info = plan.info()
    x = synth.vector_device(csr.cols, 2)
    y = synth.vector_device(csr.rows, 3)
    for _ in range(5):
        plan.execute(alpha, beta, x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.execute(alpha, beta, x, y)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    plan.destroy()
    return ms, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c2")
    ap.add_argument("--tiles", default="1024,2048,4096,8192")
    ap.add_argument("--vecdivs", default="8")
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--no-tma-too", action="store_true")
    args = ap.parse_args()
    for w in args.workloads.split(","):
        csr = make(w)
        torch.cuda.synchronize()
        balg = synth.algorithmic_bytes(csr.rows, csr.cols, csr.nnz)
        for T in [int(t) for t in args.tiles.split(",")]:
            for vd in [int(v) for v in args.vecdivs.split(",")]:
                for flags in ([0, FLAG_NO_TMA] if args.no_tma_too else [0]):
                    try:
                        ms, info = time_plan(csr, make_options(T, 0, 0, vd, flags), args.reps)
                        print(json.dumps({"workload": w, "tile": T, "vec_div": vd, "flags": flags, "ms": round(ms, 5),
                                          "gbs": round(balg / ms / 1e6, 1), "gflops": round(2 * csr.nnz / ms / 1e6, 1),
                                          "kinds": list(info.tiles_per_kind), "split": info.nsplit_rows,
                                          "launches": info.launches_per_execute, "smem": info.smem_bytes}), flush=True)
                    except Exception as e:
                        print(json.dumps({"workload": w, "tile": T, "error": str(e)}), flush=True)
        del csr
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
