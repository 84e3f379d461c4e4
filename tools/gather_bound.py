"""Gather bound of a column stream (development / evidence tool): time to stream colindex (+ value) and gather
x[colindex[k]] with no row structure at all (spmv_b200_ctx_gather_bound). Any CSR kernel that gathers x per element
pays at least this much for the same matrix. Also: the same with part of the unified L1/shared array taken away
(--smem), and with other flavours of the gather load. Prints one JSON line per case.

    python tools/gather_bound.py [--workloads c3,c4] [--tables 2500000,...] [--smem 0,16,32] [--flavours 0,1,2]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import torch  # noqa: E402

from spmv_acc_b200 import _lib, synth  # noqa: E402
from sweep import make  # noqa: E402

FLAVOURS = ["ld.global.nc", "L1::no_allocate", "ld.global.cg", "L1::evict_first", "L1::evict_last"]


def time_case(col, val, x, nnz, with_val, flavour, cps, smem_kb, reps):
    X = _lib.ctx()
    mode = (1 if with_val else 0) | (flavour << 1)
    grid = X.spmv_b200_ctx_gather_bound(nnz, 0, 0, 0, 0, mode, cps, 0, 0)
    out = torch.empty(grid * 256, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    args = (nnz, col.data_ptr(), val.data_ptr() if val is not None else 0, x.data_ptr(), out.data_ptr(), mode, cps,
            smem_kb * 1024, st)
    for _ in range(3):
        rc = X.spmv_b200_ctx_gather_bound(*args)
        assert rc == 0, rc
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        X.spmv_b200_ctx_gather_bound(*args)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def time_gather4(col, x, nnz, cps, box_rows, reps):
    """cp.async.bulk.tensor tile::gather4 flavour (spmv_b200_ctx_gather4_bound). Returns (ms, sum of the gathered x)
    or (None, error code)."""
    X = _lib.ctx()
    n = x.numel()
    grid = X.spmv_b200_ctx_gather4_bound(nnz, 0, 0, n, 0, cps, box_rows, 0)
    out = torch.zeros(grid * 256, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    args = (nnz, col.data_ptr(), x.data_ptr(), n, out.data_ptr(), cps, box_rows, st)
    rc = X.spmv_b200_ctx_gather4_bound(*args)
    if rc != 0:
        return None, rc
    torch.cuda.synchronize()
    total = float(out.sum().item())
    for _ in range(2):
        X.spmv_b200_ctx_gather4_bound(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        X.spmv_b200_ctx_gather4_bound(*args)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps, total


def gather4_cases(a):
    """TMA gather4 against the LSU gather on uniformly random columns (and C3's own column stream)."""
    nnz = 320_000_000
    cases = [("uniform-random columns", int(t)) for t in a.tables.split(",") if t]
    for name, n in cases:
        col = torch.randint(0, n, (nnz,), dtype=torch.int32, device="cuda")
        x = synth.vector_device(n, 2)
        want = float(x[col.long()].sum().item()) if n <= 20_000_000 else None
        lsu = time_case(col, None, x, nnz, False, 0, 8, 0, a.reps)
        print(json.dumps({"case": name, "nnz": nnz, "x_MB": n * 8 / 1e6, "gather": "ld.global.nc (LSU)",
                          "ms": round(lsu, 4), "Ggather_s": round(nnz / lsu / 1e6, 1)}), flush=True)
        for box_rows in (1, 4):
            for cps in (2, 3, 4):
                ms, total = time_gather4(col, x, nnz, cps, box_rows, a.reps)
                rec = {"case": name, "nnz": nnz, "x_MB": n * 8 / 1e6,
                       "gather": f"TMA tile::gather4, x as [n/2][2] fp64, box {{2,{box_rows}}}", "ctas_per_sm": cps}
                if ms is None:
                    rec["error"] = total
                else:
                    rec.update({"ms": round(ms, 4), "Ggather_s": round(nnz / ms / 1e6, 1), "sum": total,
                                "sum_expected": want,
                                "sum_ok": (abs(total - want) <= 1e-9 * max(1.0, abs(want))) if want is not None else None})
                print(json.dumps(rec), flush=True)
        del col, x
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c3,c4")
    ap.add_argument("--tables", default="1250000,2500000,5000000,7500000,10000000,20000000")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--ctas", default="8")
    ap.add_argument("--smem", default="0", help="KB of dynamic shared memory per CTA (shrinks L1), comma list")
    ap.add_argument("--flavours", default="0,1")
    ap.add_argument("--gather4", action="store_true", help="only the TMA tile::gather4 comparison")
    a = ap.parse_args()
    if a.gather4:
        gather4_cases(a)
        return
    fls = [int(f) for f in a.flavours.split(",")]
    for w in [w for w in a.workloads.split(",") if w]:
        csr = make(w)
        x = synth.vector_device(csr.cols, 2)
        balg = synth.algorithmic_bytes(csr.rows, csr.cols, csr.nnz)
        for cps in [int(c) for c in a.ctas.split(",")]:
            for kb in [int(k) for k in a.smem.split(",")]:
                for fl in fls:
                    for wv in (False, True):
                        ms = time_case(csr.col, csr.val, x, csr.nnz, wv, fl, cps, kb, a.reps)
                        print(json.dumps({"case": w, "nnz": csr.nnz, "x_MB": csr.cols * 8 / 1e6, "with_val": wv,
                                          "gather": FLAVOURS[fl], "ctas_per_sm": cps, "smem_kb_per_cta": kb,
                                          "smem_kb_per_sm": kb * cps, "ms": round(ms, 4),
                                          "Ggather_s": round(csr.nnz / ms / 1e6, 1),
                                          "spmv_equiv_gbs": round(balg / ms / 1e6, 1)}), flush=True)
        del csr, x
        torch.cuda.empty_cache()
    nnz = 320_000_000
    val = None
    for n in [int(t) for t in a.tables.split(",") if t]:
        if val is None:
            val = synth.vector_device(nnz, 5)
        col = torch.randint(0, n, (nnz,), dtype=torch.int32, device="cuda")
        x = synth.vector_device(n, 2)
        for fl in fls:
            ms = time_case(col, val, x, nnz, True, fl, 8, 0, a.reps)
            print(json.dumps({"case": "uniform-random columns", "nnz": nnz, "x_MB": n * 8 / 1e6, "with_val": True,
                              "gather": FLAVOURS[fl], "ctas_per_sm": 8, "ms": round(ms, 4),
                              "Ggather_s": round(nnz / ms / 1e6, 1)}), flush=True)
        del col, x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
