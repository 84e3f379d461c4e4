"""Fused halo push (SpMV epilogue stores boundary rows straight into the neighbour's x through a CUDA IPC mapping,
iterations ordered by stream flags): two processes share cuda:0 (gloo for the set-up plumbing only), and the
resulting x must be bitwise equal to the single-process loop and within the fp64 bound of the oracle."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, N, iters, out_dir):
    import torch
    import torch.distributed as dist
    from spmv_acc_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    loop, plan, csr = sharded.build_stencil3d_power_loop(N, "halo")
    assert loop.mode == "halo" and loop.sends and loop.recvs
    fused = sharded.FusedHaloLoop(loop, plan)
    # large enough grids take the boundary-first schedule (boundary row blocks + push, flags, then the interior)
    assert fused.split or N < 64, (fused.split, len(loop.boundary), len(loop.interior))
    fused.run(2, native=False)          # python-driven iterations and the natively enqueued loop must chain
    x = fused.run(iters - 2)
    torch.cuda.synchronize()
    lo, hi = int(loop.bounds[rank]), int(loop.bounds[rank + 1])
    np.save(Path(out_dir) / f"x_{rank}.npy", x[lo:hi].cpu().numpy())
    np.save(Path(out_dir) / f"meta_{rank}.npy", np.array([lo, hi]))
    fused.close()
    plan.destroy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 40), (3, 40), (2, 64), (3, 64)])
def test_fused_halo_push_equals_single_process(tmp_path, world, N):
    import torch
    import torch.multiprocessing as mp
    import oracle
    from spmv_acc_b200 import CsrDesc, SpmvPlan, make_options, synth, FLAG_BETA0_SKIP_Y
    iters = 6
    mp.spawn(_worker, args=(world, _free_port(), N, iters, str(tmp_path)), nprocs=world, join=True)
    n = N ** 3
    got = np.zeros(n)
    for r in range(world):
        lo, hi = np.load(tmp_path / f"meta_{r}.npy")
        got[lo:hi] = np.load(tmp_path / f"x_{r}.npy")
    # the same shards multiplied one after the other in this process (same plans, hence the same summation order per
    # row: the tile shapes, and with them the lanes per row, depend on where a shard starts)
    from spmv_acc_b200 import shard_bounds
    counts = synth.stencil_row_counts_device("stencil3d", N)
    bounds = shard_bounds(synth._rowptr_from_counts_device(counts), n, world).astype(np.int64)
    shards = []
    for r in range(world):
        d = synth.stencil3d_device(N, int(bounds[r]), int(bounds[r + 1]))
        shards.append((d, SpmvPlan(CsrDesc(d.rows, d.cols, d.nnz, d.rowptr, d.col, d.val),
                                   make_options(flags=FLAG_BETA0_SKIP_Y))))
    x, y = synth.vector_device(n, 2), torch.zeros(n, dtype=torch.float64, device="cuda")
    for _ in range(iters):
        for r, (d, plan) in enumerate(shards):
            plan.execute(1.0, 0.0, x, y[int(bounds[r]):int(bounds[r + 1])])
        x, y = y, x
    torch.cuda.synchronize()
    one = x.cpu().numpy()
    for _, plan in shards:
        plan.destroy()
    bad = np.flatnonzero(got != one)
    assert bad.size == 0, f"fused halo loop differs from the sequential shard loop at {bad.size} rows, first {bad[:4]}"
    # and the oracle, iterated on the host
    h = synth.stencil3d_numpy(N)
    xr = synth.vector_numpy(n, 2)
    for _ in range(iters):
        xr = oracle.best_host_spmv(1.0, 0.0, h.rowptr, h.col, h.val, xr, np.zeros(n))
    assert np.max(np.abs(got - xr)) <= 1e-12 * max(1.0, np.max(np.abs(xr)))
