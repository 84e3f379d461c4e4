"""Multi-rank host logic on CPU: world_size-2 gloo runs of the sharded power loop. The compute callback is the oracle
here (test-only injection; the product wires in the CUDA plan, see spmv_acc_b200/sharded.py:build_stencil3d_power_loop)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from spmv_acc_b200 import sharded, synth  # noqa: E402


def test_merge_runs_and_schedule_are_consistent():
    need = np.zeros((3, 8), np.uint8)
    bounds = np.array([0, 10000, 20000, 30000])
    need[0, [2, 3]] = 1          # rank 0 needs blocks 2,3 (elements 8192..16383): own + rank 1's 10000..16383
    need[1, [0, 2, 4, 5]] = 1
    need[2, :] = 1
    for r in range(3):
        sends, _ = sharded.exchange_schedule(need, bounds, r)
        for p, a, e in sends:
            _, recvs_p = sharded.exchange_schedule(need, bounds, p)
            assert (r, a, e) in recvs_p
    _, recvs0 = sharded.exchange_schedule(need, bounds, 0)
    assert recvs0 == [(1, 10000, 16384)]
    assert sharded.merge_runs(np.array([0, 1, 3]), 100, 13000) == [(100, 8192), (12288, 13000)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, mode, N, iters, out_dir):
    import torch
    import torch.distributed as dist
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = synth.stencil3d_numpy(N)
    n = N ** 3
    bounds = oracle.port_shard_bounds(full.rowptr, world).astype(np.int64)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    shard = synth.stencil3d_numpy(N, lo, hi)
    assert np.array_equal(shard.col, full.col[full.rowptr[lo]:full.rowptr[hi]])
    shift = 6  # 64-entry blocks so that the small test grid has a real halo structure
    need = np.zeros((n + (1 << shift) - 1) >> shift, np.uint8)
    need[np.unique(shard.col >> shift)] = 1

    def spmv(xf, ys):  # test-only compute: the oracle
        y = oracle.port_host_spmv(1.0, 0.0, shard.rowptr, shard.col, shard.val, xf.numpy(), np.zeros(hi - lo))
        ys.copy_(torch.from_numpy(y))

    # row blocks of ~97 rows: lets the halo mode compute the rows other ranks wait for first (overlap path)
    tile_row = np.unique(np.concatenate([np.arange(0, hi - lo, 97), [hi - lo]])).astype(np.int64)

    def spmv_tiles(xf, ys, t0, t1):
        a, b = int(tile_row[t0]), int(tile_row[t1])
        sub_rp = shard.rowptr[a:b + 1] - shard.rowptr[a]
        sl = slice(int(shard.rowptr[a]), int(shard.rowptr[b]))
        y = oracle.port_host_spmv(1.0, 0.0, sub_rp, shard.col[sl], shard.val[sl], xf.numpy(), np.zeros(b - a))
        ys[a:b].copy_(torch.from_numpy(y))

    # which row blocks read x entries owned by the other rank (the product gets this from spmv_b200_plan_tile_col_range)
    reads_halo = np.zeros(tile_row.size - 1, dtype=bool)
    for t in range(tile_row.size - 1):
        cols = shard.col[shard.rowptr[tile_row[t]]:shard.rowptr[tile_row[t + 1]]]
        reads_halo[t] = cols.size > 0 and (cols.min() < lo or cols.max() >= hi)
    x = torch.from_numpy(synth.vector_numpy(n, 2).copy())
    loop = sharded.PowerLoop(n=n, bounds=bounds, spmv=spmv, x=x, x_next=torch.zeros_like(x), need_local=need,
                             exchange=mode, block_shift=shift, spmv_tiles=spmv_tiles, tile_row=tile_row,
                             tile_reads_halo=reads_halo)
    if mode == "halo":
        assert loop.overlapped and loop.boundary and loop.interior and loop.boundary_reads_all_halo
        covered = np.zeros(reads_halo.size, dtype=bool)
        for t0, t1 in loop.boundary:
            covered[t0:t1] = True
        assert covered[reads_halo].all()      # every reader of halo entries is a boundary row block
        assert not covered.all()
    xf = loop.run(iters)
    np.save(Path(out_dir) / f"x_{mode}_{rank}.npy", xf[lo:hi].numpy())
    np.save(Path(out_dir) / f"meta_{mode}_{rank}.npy", np.array([lo, hi, loop.bytes_in_per_iter]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["allgather", "halo"])
def test_two_rank_power_loop_equals_single_process(tmp_path, mode):
    import torch.multiprocessing as mp
    import oracle
    N, iters, world = 20, 4, 2
    mp.spawn(_worker, args=(world, _free_port(), mode, N, iters, str(tmp_path)), nprocs=world, join=True)
    full = synth.stencil3d_numpy(N)
    x = synth.vector_numpy(N ** 3, 2)
    for _ in range(iters):
        x = oracle.port_host_spmv(1.0, 0.0, full.rowptr, full.col, full.val, x, np.zeros(N ** 3))
    got = np.zeros_like(x)
    moved = []
    for r in range(world):
        lo, hi, nbytes = np.load(tmp_path / f"meta_{mode}_{r}.npy")
        got[lo:hi] = np.load(tmp_path / f"x_{mode}_{r}.npy")
        moved.append(int(nbytes))
    assert np.array_equal(got, x)  # same arithmetic per row -> bitwise equal to the single-process loop
    if mode == "halo":
        # z-slab shards of a 27-point stencil need about one plane per neighbour, not the whole vector
        assert max(moved) < 0.5 * 8 * N ** 3 / 2


def test_align_bounds_rounds_interior_boundaries_to_lines_of_x():
    from spmv_acc_b200 import sharded
    b = np.array([0, 7114817, 14180438, 21246059, 56623104], dtype=np.int64)
    a = sharded.align_bounds(b)
    assert a[0] == 0 and a[-1] == b[-1] and np.all(a[1:-1] % 16 == 0) and np.all(np.abs(a - b) <= 8)
    assert np.all(np.diff(a) >= 0)
    # degenerate inputs: one shard, empty shards, tiny matrices
    assert sharded.align_bounds(np.array([0, 100])).tolist() == [0, 100]
    assert sharded.align_bounds(np.array([0, 3, 3, 5])).tolist() == [0, 0, 0, 5]
    s, r = sharded.allgather_schedule(a, 1)
    assert [p for p, _, _ in s] == [0, 2, 3] and all((lo, hi) == (a[1], a[2]) for _, lo, hi in s)
    assert [(p, lo, hi) for p, lo, hi in r] == [(0, a[0], a[1]), (2, a[2], a[3]), (3, a[3], a[4])]
