"""Where does the time of the all-gather mode of the iterated loop go? (run under torchrun, development tool)
Variants: exchange alone / SpMV alone / both serialised / overlapped, eager and as a CUDA graph.

    torchrun --nproc-per-node 8 tools/allgather_loop_probe.py [N] [iters]
"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from spmv_acc_b200 import sharded  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 384
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")


def p2p_allgather(v, bounds):
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    ops = []
    for d in range(1, world):
        dst, src = (rank + d) % world, (rank - d) % world
        ops.append(dist.P2POp(dist.isend, v[lo:hi], dst))
        ops.append(dist.P2POp(dist.irecv, v[int(bounds[src]):int(bounds[src + 1])], src))
    for r in dist.batch_isend_irecv(ops):
        r.wait()


shard = None


def timed(fn, reps):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 4)


def report(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


import numpy as np  # noqa: E402

n = N ** 3
eq = np.array([(n * g) // world for g in range(world + 1)], dtype=np.int64)
fresh = torch.zeros(n, dtype=torch.float64, device="cuda")
report(what="fresh process: grouped point-to-point all-gather of a fresh tensor, equal slices",
       ms=timed(lambda: p2p_allgather(fresh, eq), iters))
shard = sharded.build_shard("stencil3d", N=N)
report(what="after build_shard: the same call on the same tensor", ms=timed(lambda: p2p_allgather(fresh, eq), iters),
       bounds=[int(b) for b in shard.bounds])
report(what="after build_shard: the same with the shard's nnz-balanced bounds",
       ms=timed(lambda: p2p_allgather(fresh, shard.bounds), iters))
torch.cuda.empty_cache()
free, total = torch.cuda.mem_get_info()
report(what="device memory", free_GB=round(free / 1e9, 2), total_GB=round(total / 1e9, 2))
for overlap in (False, True):
    loop = sharded.make_loop(shard, "allgather", overlap=overlap)
    lo, hi = int(loop.bounds[rank]), int(loop.bounds[rank + 1])
    report(what="exchange alone (loop._allgather on the current stream)", overlap=overlap,
           ms=timed(lambda: loop._allgather(loop.x), iters))
    report(what="SpMV alone (whole shard)", overlap=overlap,
           ms=timed(lambda: loop.spmv(loop.x, loop.x_next[lo:hi]), iters))
    if overlap:
        report(what="interior row blocks alone", n=int(sum(b - a for a, b in loop.interior)),
               ms=timed(lambda: [loop.spmv_tiles(loop.x, loop.x_next[lo:hi], a, b) for a, b in loop.interior], iters))
        report(what="boundary row blocks alone", n=int(sum(b - a for a, b in loop.boundary)),
               ms=timed(lambda: [loop.spmv_tiles(loop.x, loop.x_next[lo:hi], a, b) for a, b in loop.boundary], iters))
    for sms in ((0, 32) if overlap else (0,)):
        shard.plan.set_comm_sms(sms)
        report(what="loop.step eager", overlap=overlap, comm_sms=sms, ms=timed(loop.step, iters))
        try:
            loop.capture()
            report(what="loop.step as CUDA graph (2 iterations per replay)", overlap=overlap, comm_sms=sms,
                   ms=round(timed(loop.graph.replay, iters // 2) / 2, 4))
        except Exception as e:
            report(what="graph", error=f"{type(e).__name__}: {e}"[:200])
        shard.plan.set_comm_sms(0)
    del loop
dist.destroy_process_group()
