"""Fused halo loop (spmv_b200_halo_loop_*): the SpMV kernels store the rows a neighbour references straight into that
neighbour's copy of the next x and order the iterations with flags they wait for and raise themselves.

One GPU is enough to check the protocol without ever running two kernels that wait for each other (which nothing
guarantees to be co-scheduled on one device): the shards of all "ranks" live in this process, and their iterations are
enqueued round-robin on one stream, so every flag a kernel waits for has already been raised when it starts. What is
checked: bitwise equality with the same shards multiplied one after the other, the fp64 bound against the oracle
iterated on the host, both launch forms (one launch per iteration / wait + boundary + flag + interior launches), the
graph replay (single rank), and that a neighbour that never answers turns into SPMV_B200_ERR_TIMEOUT instead of a wrong x.
The real multi-GPU run (peer memory over NVLink, CUDA IPC) is bench.py under torchrun, which compares the checksum of x
between 1, 2, 4 and 8 GPUs."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


def _build_ranks(N, world, flags, allgather=False):
    import torch
    from spmv_acc_b200 import (CsrDesc, HaloLoop, SpmvPlan, col_block_bitmap, make_options, shard_bounds, sharded,
                               synth, FLAG_BETA0_SKIP_Y)
    n = N ** 3
    counts = synth.stencil_row_counts_device("stencil3d", N)
    bounds = shard_bounds(synth._rowptr_from_counts_device(counts), n, world).astype(np.int64)
    shift = sharded.BLOCK_SHIFT
    shards, plans, need = [], [], []
    for r in range(world):
        d = synth.stencil3d_device(N, int(bounds[r]), int(bounds[r + 1]))
        shards.append(d)
        plans.append(SpmvPlan(CsrDesc(d.rows, d.cols, d.nnz, d.rowptr, d.col, d.val), make_options(flags=FLAG_BETA0_SKIP_Y)))
        need.append(col_block_bitmap(d.col, d.nnz, n, shift))
    need = np.stack(need)
    x0 = synth.vector_device(n, 2)
    bufs = [[x0.clone(), torch.zeros_like(x0)] for _ in range(world)]
    flag_words = [torch.zeros(world, dtype=torch.int32, device="cuda") for _ in range(world)]
    loops, splits, all_recvs = [], [], []
    for r in range(world):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        sends, recvs = (sharded.allgather_schedule(bounds, r) if allgather else
                        sharded.exchange_schedule(need, bounds, r, shift))
        all_recvs.append(recvs)
        neigh = sorted({p for p, _, _ in sends} | {p for p, _, _ in recvs})
        cmin, cmax = plans[r].tile_col_range()
        split = None if allgather else sharded.split_boundary_interior(sends, plans[r].export("tile_row"),
                                                                       (cmin < lo) | (cmax >= hi), lo)
        splits.append(split is not None)
        push = [[(a - lo, e - lo, bufs[p][b].data_ptr() + 8 * lo) for p, a, e in sends] for b in (0, 1)]
        desc = sharded.make_halo_desc(
            plans[r], [bufs[r][0].data_ptr(), bufs[r][1].data_ptr()], lo, hi,
            [flag_words[r].data_ptr() + 4 * p for p in neigh], [flag_words[p].data_ptr() + 4 * r for p in neigh], push,
            split[0] if split else [], split[1] if split else [], flags)
        loops.append(HaloLoop(desc))
    return dict(n=n, bounds=bounds, shards=shards, plans=plans, bufs=bufs, loops=loops, x0=x0, splits=splits,
                flag_words=flag_words, recvs=all_recvs)


@pytest.mark.parametrize("world,N,flags", [(2, 40, 0), (3, 40, 0), (2, 64, 0), (3, 64, 0), (3, 64, 3), (2, 40, 3),
                                           (3, 64, 4), (2, 40, 7)])  # 4 = SPMV_B200_HALO_ALIGN_PUSH
def test_fused_halo_loop_equals_sequential_shards_and_oracle(world, N, flags):
    import torch
    import oracle
    from spmv_acc_b200 import synth
    R = _build_ranks(N, world, flags)
    n, bounds, iters = R["n"], R["bounds"], 6
    assert any(R["splits"]) or N < 64, "large grids take the boundary-first schedule"
    for lp, plan in zip(R["loops"], R["plans"]):
        i, kinds = lp.info(), list(plan.info().tiles_per_kind)
        one_row_kind = kinds[2] == 0 and (kinds[0] == 0 or kinds[1] == 0)
        if (flags & 3) == 0 and one_row_kind:  # one launch per iteration, flag protocol inside the kernel
            assert i.single_launch == 1 and i.launches_per_iteration == 1
        else:
            assert i.single_launch == 0 and i.launches_per_iteration >= 3
    for _ in range(iters):          # round-robin: every awaited flag has been raised by an earlier launch
        for lp in R["loops"]:
            lp.run(1)
    for lp in R["loops"]:
        lp.sync()
    got = torch.empty(n, dtype=torch.float64, device="cuda")
    for r in range(world):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        got[lo:hi] = R["bufs"][r][iters % 2][lo:hi]
    # every rank's flag has reached `iters` in every neighbour's word
    for r in range(world):
        fw = R["flag_words"][r].cpu().numpy()
        assert set(fw[fw > 0].tolist()) <= {iters}
    # the same shards multiplied one after the other (same plans, hence the same summation order per row)
    x, y = R["x0"].clone(), torch.zeros(n, dtype=torch.float64, device="cuda")
    for _ in range(iters):
        for r, plan in enumerate(R["plans"]):
            plan.execute(1.0, 0.0, x, y[int(bounds[r]):int(bounds[r + 1])])
        x, y = y, x
    torch.cuda.synchronize()
    bad = torch.nonzero(got != x).flatten()
    assert bad.numel() == 0, f"fused halo loop differs from the sequential shard loop at {bad.numel()} rows"
    # halo entries a rank received equal the owner's values (the push reached the right place)
    for r in range(world):
        buf = R["bufs"][r][iters % 2]
        assert R["recvs"][r], "every rank of a z-slab sharded stencil has a halo"
        for _, a, e in R["recvs"][r]:
            assert torch.equal(buf[a:e], x[a:e])
    h = synth.stencil3d_numpy(N)
    xr = synth.vector_numpy(n, 2)
    for _ in range(iters):
        xr = oracle.best_host_spmv(1.0, 0.0, h.rowptr, h.col, h.val, xr, np.zeros(n))
    assert np.max(np.abs(got.cpu().numpy() - xr)) <= 1e-12 * max(1.0, np.max(np.abs(xr)))
    for lp in R["loops"]:
        lp.destroy()
    for p in R["plans"]:
        p.destroy()


@pytest.mark.parametrize("world,N,flags", [(3, 40, 0), (4, 48, 0), (3, 40, 3), (3, 40, 4), (4, 48, 4), (3, 48, 7)])
def test_fused_allgather_push_leaves_the_whole_x_everywhere(world, N, flags):
    """The all-gather done by the SpMV kernels: every rank pushes its whole slice to every other rank (whole-shard
    schedule: all peers' flags are awaited before the first row block, the own flag is raised after the last). After k
    iterations EVERY rank's buffer must hold the complete x_k, bitwise equal to the sequential shard loop."""
    import torch
    R = _build_ranks(N, world, flags, allgather=True)
    n, bounds, iters = R["n"], R["bounds"], 5
    for _ in range(iters):
        for lp in R["loops"]:
            lp.run(1)
    for lp in R["loops"]:
        lp.sync()
    x, y = R["x0"].clone(), torch.zeros(n, dtype=torch.float64, device="cuda")
    for _ in range(iters):
        for r, plan in enumerate(R["plans"]):
            plan.execute(1.0, 0.0, x, y[int(bounds[r]):int(bounds[r + 1])])
        x, y = y, x
    torch.cuda.synchronize()
    for r in range(world):
        assert torch.equal(R["bufs"][r][iters % 2], x), f"rank {r} does not hold the whole x after the fused all-gather"
        fw = R["flag_words"][r].cpu().numpy()
        assert sorted(np.nonzero(fw)[0].tolist()) == [p for p in range(world) if p != r] and set(fw[fw > 0]) == {iters}
    for lp in R["loops"]:
        lp.destroy()
    for p in R["plans"]:
        p.destroy()


@pytest.mark.parametrize("flags", [0, 2])
def test_single_rank_loop_replays_graph_and_matches_eager(flags, monkeypatch):
    """No neighbours: runs of iterations are replayed from the CUDA graph (chunks of 4 here), odd starts and remainders
    are enqueued one by one; x must be bitwise equal to plan.execute in a python loop."""
    import torch
    from spmv_acc_b200 import CsrDesc, HaloLoop, SpmvPlan, make_options, sharded, synth, FLAG_BETA0_SKIP_Y
    monkeypatch.setenv("SPMV_B200_HALO_GRAPH_CHUNK", "4")
    N = 48
    n = N ** 3
    d = synth.stencil3d_device(N)
    plan = SpmvPlan(CsrDesc(d.rows, d.cols, d.nnz, d.rowptr, d.col, d.val), make_options(flags=FLAG_BETA0_SKIP_Y))
    x0 = synth.vector_device(n, 2)
    bufs = [x0.clone(), torch.zeros_like(x0)]
    lp = HaloLoop(sharded.make_halo_desc(plan, [bufs[0].data_ptr(), bufs[1].data_ptr()], 0, n, [], [], [[], []], [], [],
                                         flags))
    info = lp.info()
    assert info.uses_graph == 4 and info.single_launch == (1 if flags == 0 else 0)
    lp.run(3)      # eager, ends on an odd iteration
    lp.run(14)     # 1 eager + 3 graph launches + 1 eager
    lp.sync()
    assert lp.info().iterations_enqueued == 17
    x, y = x0.clone(), torch.zeros_like(x0)
    for _ in range(17):
        plan.execute(1.0, 0.0, x, y)
        x, y = y, x
    torch.cuda.synchronize()
    assert torch.equal(bufs[17 % 2], x)
    lp.destroy()
    plan.destroy()


@pytest.mark.parametrize("flags", [0, 3])
def test_silent_neighbour_is_reported_as_timeout(flags, monkeypatch):
    """A neighbour whose flag never advances: the second iteration gives up after SPMV_B200_FLAG_TIMEOUT_MS, does not
    raise its own flag again, and halo_loop_sync reports SPMV_B200_ERR_TIMEOUT (never a silently stale x)."""
    import torch
    from spmv_acc_b200 import CsrDesc, HaloLoop, SpmvB200Error, SpmvPlan, _lib, make_options, sharded, synth, FLAG_BETA0_SKIP_Y
    monkeypatch.setenv("SPMV_B200_FLAG_TIMEOUT_MS", "200")
    N = 32
    n = N ** 3
    half = n // 2
    d = synth.stencil3d_device(N, 0, half)
    plan = SpmvPlan(CsrDesc(d.rows, d.cols, d.nnz, d.rowptr, d.col, d.val), make_options(flags=FLAG_BETA0_SKIP_Y))
    bufs = [synth.vector_device(n, 2), torch.zeros(n, dtype=torch.float64, device="cuda")]
    mine = torch.zeros(1, dtype=torch.int32, device="cuda")      # written by the (absent) neighbour: stays 0
    theirs = torch.zeros(1, dtype=torch.int32, device="cuda")    # written by this rank
    lp = HaloLoop(sharded.make_halo_desc(plan, [bufs[0].data_ptr(), bufs[1].data_ptr()], 0, half, [mine.data_ptr()],
                                         [theirs.data_ptr()], [[], []], [], [], flags))
    lp.run(1)
    lp.sync()                                   # iteration 0 waits for nothing
    assert int(theirs.item()) == 1
    lp.run(2)                                   # iteration 1 waits for a flag that never comes
    with pytest.raises(SpmvB200Error) as err:
        lp.sync()
    assert err.value.args and "in time" in str(err.value)
    assert int(theirs.item()) == 1              # no signal after the time-out: the failure travels, stale rows do not
    lp.destroy()
    plan.destroy()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ipc_worker(rank, world, port, out_dir):
    """Two processes on cuda:0: rank 1 pushes rows of an SpMV into rank 0's exported buffer through a CUDA IPC mapping;
    ordering is by host barriers only (no kernel waits for another process' kernel)."""
    import torch
    import torch.distributed as dist
    from spmv_acc_b200 import CsrDesc, PeerBuffer, SpmvPlan, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    NB = 1 << 16
    own = PeerBuffer.alloc(8 * NB)
    buf = own.tensor("float64", NB)
    buf.zero_()
    torch.cuda.synchronize()
    objs = [None] * world
    dist.all_gather_object(objs, (own.handle, own.nbytes))
    h = synth.stencil2d_numpy(64)
    d = synth.to_device(h)
    plan = SpmvPlan(CsrDesc(d.rows, d.cols, d.nnz, d.rowptr, d.col, d.val))
    x = synth.vector_device(d.cols, 2)
    y = torch.zeros(d.rows, dtype=torch.float64, device="cuda")
    pb = None
    if rank == 1:
        pb = PeerBuffer.open(*objs[0])
        plan.execute_push(1.0, 0.0, x, y, [(10, 50, pb.address + 8 * 1000)])
        torch.cuda.synchronize()
    else:
        plan.execute(1.0, 0.0, x, y)
        torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        ok = bool(torch.equal(buf[1010:1050], y[10:50])) and float(buf[:1010].abs().sum()) == 0.0
        np.save(Path(out_dir) / "ok.npy", np.array([int(ok)]))
    dist.barrier()
    if pb is not None:
        pb.release()
    dist.barrier()
    own.release()
    plan.destroy()
    dist.destroy_process_group()


def test_push_into_another_process_buffer_through_cuda_ipc(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_ipc_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert int(np.load(tmp_path / "ok.npy")[0]) == 1
