"""Aggregate pinned-memory copy bandwidth of the box with all GPUs copying at once (run under torchrun): what bounds
the host-buffer path (e2e) at N > 1. Every rank copies `mb` MB host->device, device->host, and both at once, between
two barriers; rank 0 prints the sum over ranks.

    torchrun --nproc-per-node 8 tools/pcie_aggregate_probe.py [mb]
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 57
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
n = mb * (1 << 20) // 8
h_in = torch.ones(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


for name, fn, factor in (("h2d", h2d, 1), ("d2h", d2h, 1), ("both directions at once", both, 2)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / 20
    t = torch.tensor([sec], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        per_rank = factor * n * 8 / float(t.item()) / 1e9
        print(json.dumps({"copy": name, "world": world, "MB_per_rank_and_direction": mb, "ms": round(float(t.item()) * 1e3, 3),
                          "GBs_per_rank": round(per_rank, 1), "GBs_all_ranks": round(per_rank * world, 1)}), flush=True)
dist.destroy_process_group()
