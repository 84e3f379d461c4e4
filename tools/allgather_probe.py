"""All-gather of unequal slices of x over NCCL, four ways (development / evidence tool, run under torchrun):
    (a) dist.all_gather with a list of unequal tensors (ProcessGroupNCCL: a group of broadcasts)
    (b) grouped point-to-point send / recv, in place
    (c) ncclAllGather of equal padded slots into a staging buffer + device copies into place
    (d) ncclAllGather in place on equal slices (what (a)-(c) are measured against)
Prints one JSON line per method on rank 0.

    torchrun --nproc-per-node 8 tools/allgather_probe.py [n] [reps]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384 ** 3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
# nnz-balanced slabs of a 27-point stencil: the two edge slabs are a little longer
base = n // world
skew = base // 100
sizes = np.full(world, base, dtype=np.int64)
if world > 2:
    sizes[0] += skew
    sizes[-1] += skew
    sizes[1:-1] -= (2 * skew) // (world - 2)
sizes[-1] += n - sizes.sum()
bounds = np.concatenate([[0], np.cumsum(sizes)])
lo, hi = int(bounds[rank]), int(bounds[rank + 1])
x = torch.zeros(n, dtype=torch.float64, device="cuda")
x[lo:hi] = rank + 1.0
maxlen = int(sizes.max())
staging = torch.empty(world * maxlen, dtype=torch.float64, device="cuda")
padded = torch.zeros(maxlen, dtype=torch.float64, device="cuda")


def a_bcast_group():
    dist.all_gather([x[int(bounds[g]):int(bounds[g + 1])] for g in range(world)], x[lo:hi])


def b_p2p():
    ops = []
    for d in range(1, world):
        dst, src = (rank + d) % world, (rank - d) % world
        ops.append(dist.P2POp(dist.isend, x[lo:hi], dst))
        ops.append(dist.P2POp(dist.irecv, x[int(bounds[src]):int(bounds[src + 1])], src))
    for r in dist.batch_isend_irecv(ops):
        r.wait()


def c_padded():
    padded[:hi - lo].copy_(x[lo:hi])
    dist.all_gather_into_tensor(staging, padded)
    for g in range(world):
        if g != rank:
            a, e = int(bounds[g]), int(bounds[g + 1])
            x[a:e].copy_(staging[g * maxlen:g * maxlen + (e - a)])


xe = torch.zeros(world * base, dtype=torch.float64, device="cuda")


def d_equal_in_place():
    dist.all_gather_into_tensor(xe, xe[rank * base:(rank + 1) * base])


def check():
    want = sum((g + 1.0) * float(sizes[g]) for g in range(world))
    return abs(float(x.sum().item()) - want) < 1e-6 * want


for name, fn in (("a_bcast_group", a_bcast_group), ("b_p2p", b_p2p), ("c_padded_allgather_plus_copies", c_padded),
                 ("d_equal_in_place", d_equal_in_place)):
    try:
        x.zero_()
        x[lo:hi] = rank + 1.0
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ok = True if name.startswith("d_") else check()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            inbound = 8.0 * (n - (hi - lo))
            print(json.dumps({"method": name, "world": world, "n": n, "ms": round(float(ms.item()), 4), "ok": ok,
                              "inbound_GBs_per_gpu": round(inbound / float(ms.item()) / 1e6, 1)}), flush=True)
    except Exception as e:
        if rank == 0:
            print(json.dumps({"method": name, "error": f"{type(e).__name__}: {e}"[:300]}), flush=True)
dist.destroy_process_group()
