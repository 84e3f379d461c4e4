"""Does the NCCL all-gather overlap with the SpMV of the rows that read only local columns? (run under torchrun)
Variants of one overlapped iteration, eager launches: default stream / side stream for the SpMV, launch order, SMs left
to the collective, persistent ring kernels / one row block per CTA.

    torchrun --nproc-per-node 2 tools/overlap_probe.py [N] [iters]
"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from spmv_acc_b200 import FLAG_BETA0_SKIP_Y, make_options, sharded  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 384
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")


def timed(fn, reps):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream()
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 4)


def report(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


for label, flags in (("persistent ring kernels", FLAG_BETA0_SKIP_Y), ("one row block per CTA (no ring)", FLAG_BETA0_SKIP_Y | (1 << 25))):
    shard = sharded.build_shard("stencil3d", N=N, options=make_options(flags=flags))
    loop = sharded.make_loop(shard, "allgather", overlap=True)
    lo, hi = int(loop.bounds[rank]), int(loop.bounds[rank + 1])
    ys = loop.x_next[lo:hi]
    comm = torch.cuda.Stream(priority=-1)

    def exchange():
        loop._allgather(loop.x)

    def interior():
        for a, b in loop.interior:
            loop.spmv_tiles(loop.x, ys, a, b)

    def boundary():
        for a, b in loop.boundary:
            loop.spmv_tiles(loop.x, ys, a, b)

    def step(order):
        cur = torch.cuda.current_stream()
        comm.wait_event(cur.record_event())
        if order == "exchange first":
            with torch.cuda.stream(comm):
                exchange()
                done = comm.record_event()
            interior()
        else:
            interior()
            with torch.cuda.stream(comm):
                exchange()
                done = comm.record_event()
        cur.wait_event(done)
        boundary()

    report(kernels=label, what="exchange alone", ms=timed(exchange, iters))
    report(kernels=label, what="interior + boundary alone", ms=timed(lambda: (interior(), boundary()), iters))
    for sms in (0, 32, 64):
        shard.plan.set_comm_sms(sms)
        for order in ("exchange first", "interior first"):
            report(kernels=label, what="overlapped step, SpMV on the default stream", order=order, comm_sms=sms,
                   ms=timed(lambda: step(order), iters))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                report(kernels=label, what="overlapped step, SpMV on a side stream", order=order, comm_sms=sms,
                       ms=timed(lambda: step(order), iters))
            torch.cuda.current_stream().wait_stream(side)
    shard.plan.set_comm_sms(0)
    del loop
    shard.destroy()
    torch.cuda.empty_cache()
dist.destroy_process_group()
