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
    if workload == "c3half":  # one column half of C3 (what a 2-way column-blocked plan would multiply per pass)
        return synth.uniform_device(10_000_000, 5_000_000, 16, seed=1)
    if workload == "c3x40":   # C3 with x small enough for L2 (40 MB)
        return synth.uniform_device(10_000_000, 5_000_000, 32, seed=1)
    if workload == "c4":
        return synth.rmat_device(24, 16, seed=1)
    if workload == "c5":
        return synth.stencil3d_device(384)
    if workload == "c5s":
        return synth.stencil3d_device(256)
    if workload.startswith("ss:"):  # SuiteSparse-shaped stand-in (synth.SUITESPARSE_SHAPES)
        return synth.suitesparse_like_device(workload[3:])
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


def cusparse_ms(csr, reps, balg):
    import ctypes as C
    from spmv_acc_b200 import _lib
    X = _lib.ctx()
    x = synth.vector_device(csr.cols, 2)
    y = synth.vector_device(csr.rows, 3)
    out = {}
    for alg, name in ((0, "cusparse_default"), (2, "cusparse_alg2")):
        h = C.c_void_p()
        rc = X.spmv_b200_ctx_cusparse_create(C.byref(h), csr.rows, csr.cols, csr.nnz, csr.rowptr.data_ptr(),
                                             csr.col.data_ptr(), csr.val.data_ptr(), x.data_ptr(), y.data_ptr(), alg)
        if rc:
            out[name] = f"error {rc}"
            continue
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(5):
            X.spmv_b200_ctx_cusparse_spmv(h, 1.0, 1.0, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            X.spmv_b200_ctx_cusparse_spmv(h, 1.0, 1.0, st)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = {"ms": round(ms, 5), "gbs": round(balg / ms / 1e6, 1)}
        X.spmv_b200_ctx_cusparse_destroy(h)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c2")
    ap.add_argument("--tiles", default="1024,2048,4096,8192")
    ap.add_argument("--vecdivs", default="0")
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--no-tma-too", action="store_true")
    ap.add_argument("--xflags", default="0", help="extra option flags to OR in, comma list (4 = L2 persist x)")
    ap.add_argument("--cusparse", action="store_true")
    ap.add_argument("--svar", default="0", help="SHORT kernel variants (option flag bits 8-11)")
    ap.add_argument("--mvar", default="0", help="MEDIUM kernel variants (option flag bits 12-15)")
    ap.add_argument("--ring", default="", help="staged-x ring geometries CTASxSTAGES (0 = automatic), comma list, e.g. 2x0,1x0,2x2")
    args = ap.parse_args()
    import os
    rings = [r for r in args.ring.split(",") if r] or [""]
    for w in args.workloads.split(","):
        csr = make(w)
        torch.cuda.synchronize()
        balg = synth.algorithmic_bytes(csr.rows, csr.cols, csr.nnz)
        for T in [int(t) for t in args.tiles.split(",")]:
            for vd in [int(v) for v in args.vecdivs.split(",")]:
                variants = [(int(a) << 8) | (int(b) << 12) for a in args.svar.split(",") for b in args.mvar.split(",")]
                for flags in [v | f | int(xf) for v in variants for f in ([0, FLAG_NO_TMA] if args.no_tma_too else [0])
                              for xf in args.xflags.split(",")]:
                  for ring in rings:
                    if ring:
                        c, st = ring.split("x")
                        os.environ["SPMV_B200_RING_CTAS"] = c
                        os.environ["SPMV_B200_RING_STAGES"] = st
                    try:
                        ms, info = time_plan(csr, make_options(T, 0, 0, vd, flags), args.reps)
                        print(json.dumps({"workload": w, "tile": info.tile_nnz, "vec_div": vd, "flags": flags, "ring": ring, "ms": round(ms, 5),
                                          "gbs": round(balg / ms / 1e6, 1), "gflops": round(2 * csr.nnz / ms / 1e6, 1),
                                          "kinds": list(info.tiles_per_kind), "split": info.nsplit_rows, "direct": info.direct,
                                          "xstage": info.xstage, "ring_used": f"{info.ring_ctas}x{info.ring_stages}", "launches": info.launches_per_execute, "smem": info.smem_bytes}), flush=True)
                    except Exception as e:
                        print(json.dumps({"workload": w, "tile": T, "flags": flags, "ring": ring, "error": str(e)}), flush=True)
        if args.cusparse:
            print(json.dumps({"workload": w, **cusparse_ms(csr, args.reps, balg)}), flush=True)
        del csr
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
