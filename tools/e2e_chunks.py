"""Host-buffer path (spmv_b200_hostmat_spmv) on C2 against the number of row chunks (development tool)."""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from spmv_acc_b200 import HostMatrix, synth  # noqa: E402

csr = synth.stencil2d_device(4096)
h = synth.to_host(csr)
x = synth.vector_device(csr.cols, 2)
hx = torch.empty(csr.cols, dtype=torch.float64, pin_memory=True)
hy = torch.zeros(h.rows, dtype=torch.float64, pin_memory=True)
hx.copy_(x)
hm = HostMatrix(h.rows, h.cols, h.rowptr, h.col, h.val)
for chunks in [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "8,16,32,64,4").split(",")]:
    os.environ["SPMV_B200_HOST_CHUNKS"] = str(chunks)
    for _ in range(3):
        hm.spmv(1.0, 1.0, hx, hy)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(15):
        hm.spmv(1.0, 1.0, hx, hy)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 15 * 1e3
    print(f"chunks={chunks:3d}  {ms:.3f} ms/step  {2.0 * h.nnz / ms / 1e6:.1f} GFLOP/s", flush=True)
hm.destroy()
