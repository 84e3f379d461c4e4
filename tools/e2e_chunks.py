"""Host-buffer path (spmv_b200_hostmat_spmv) against the number of row chunks, next to the raw PCIe copy rates of the
box (development / evidence tool).

    python tools/e2e_chunks.py [8,16,32,64,4] [c5|c5s|c2]
"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from spmv_acc_b200 import FLAG_BETA0_SKIP_Y, HostMatrix, make_options, synth  # noqa: E402

which = sys.argv[2] if len(sys.argv) > 2 else "c5"
csr = {"c5": lambda: synth.stencil3d_device(384), "c5s": lambda: synth.stencil3d_device(256),
       "c2": lambda: synth.stencil2d_device(4096)}[which]()
x = synth.vector_device(csr.cols, 2)
hx = torch.empty(csr.cols, dtype=torch.float64, pin_memory=True)
hy = torch.zeros(csr.rows, dtype=torch.float64, pin_memory=True)
hx.copy_(x)

# raw copy rates: one direction alone, both directions at once (two streams)
dy = torch.empty(csr.rows, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        x.copy_(hx, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hy.copy_(dy, non_blocking=True)


def both():
    h2d()
    d2h()


nb = 8.0 * csr.cols
t_in, t_out, t_both = timed(h2d), timed(d2h), timed(both)
print(json.dumps({"workload": which, "x_MB": nb / 1e6, "h2d_alone_GBs": round(nb / t_in / 1e9, 1),
                  "d2h_alone_GBs": round(nb / t_out / 1e9, 1), "both_at_once_ms": round(t_both * 1e3, 3),
                  "both_at_once_GBs_each": round(nb / t_both / 1e9, 1)}), flush=True)
del dy

hm = HostMatrix(csr.rows, csr.cols, csr.rowptr, csr.col, csr.val, make_options(flags=FLAG_BETA0_SKIP_Y))
for chunks in [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "8,16,32,64,4").split(",")]:
    os.environ["SPMV_B200_HOST_CHUNKS"] = str(chunks)
    for _ in range(3):
        hm.spmv(1.0, 0.0, hx, hy)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        hm.spmv(1.0, 0.0, hx, hy)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    print(json.dumps({"workload": which, "chunks": chunks, "ms_per_step": round(ms, 3),
                      "gflops": round(2.0 * csr.nnz / ms / 1e6, 1),
                      "floor_both_directions_ms": round(t_both * 1e3, 3)}), flush=True)
hm.destroy()
