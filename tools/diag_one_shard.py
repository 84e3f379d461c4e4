"""Builds shard `g` of `G` of the 27-point N^3 stencil on one GPU and times its SpMV for several tile sizes."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from spmv_acc_b200 import CsrDesc, SpmvPlan, make_options, shard_bounds, synth, FLAG_BETA0_SKIP_Y  # noqa: E402

N, G, g = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = N ** 3
counts = synth.stencil_row_counts_device("stencil3d", N)
bounds = shard_bounds(synth._rowptr_from_counts_device(counts), n, G).astype(np.int64)
del counts
lo, hi = int(bounds[g]), int(bounds[g + 1])
csr = synth.stencil3d_device(N, lo, hi)
x = synth.vector_device(n, 2)
y = torch.zeros(hi - lo, dtype=torch.float64, device="cuda")
for T in [0, 3072, 3328, 3584]:
    plan = SpmvPlan(CsrDesc(csr.rows, csr.cols, csr.nnz, csr.rowptr, csr.col, csr.val),
                    make_options(T, flags=FLAG_BETA0_SKIP_Y))
    info = plan.info()
    tr = plan.export("tile_row").astype(np.int64)
    rows = np.diff(tr)
    for _ in range(5):
        plan.execute(1.0, 0.0, x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        plan.execute(1.0, 0.0, x, y)
    e1.record()
    e1.synchronize()
    print(f"shard {g}/{G}: requested T={T} -> T={info.tile_nnz} tiles={info.ntiles} kinds={list(info.tiles_per_kind)} "
          f"rows/tile max={rows.max()} mean={rows.mean():.2f} over128={(rows > 128).mean():.3f} "
          f"avg nnz/row={csr.nnz / csr.rows:.4f}  {e0.elapsed_time(e1) / 50:.4f} ms", flush=True)
    plan.destroy()
