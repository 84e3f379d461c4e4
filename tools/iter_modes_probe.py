"""NCCL-exchange modes of the iterated loop against the number of SMs left to the collective (run under torchrun).

    torchrun --nproc-per-node 8 tools/iter_modes_probe.py [N] [iters]
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
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
shard = sharded.build_shard("stencil3d", N=N)
for mode in ("nccl_allgather", "nccl_halo"):
    for sms in (0, 16, 32, 48, 64):
        os.environ["SPMV_B200_COMM_SMS"] = str(sms)
        try:
            r = sharded.time_power_loop(shard, mode, iters=iters, warmup=6)
            if rank == 0:
                print(json.dumps({"mode": mode, "sms_left_to_the_collective": sms, "world": world,
                                  "ms_per_iter": round(r["ms_per_iter"], 4), "checksum": r["x_checksum_first_16th"],
                                  "overlapped": r["exchange_overlapped_with_local_rows"]}), flush=True)
        except Exception as e:
            if rank == 0:
                print(json.dumps({"mode": mode, "sms": sms, "error": f"{type(e).__name__}: {e}"[:300]}), flush=True)
r = sharded.time_power_loop(shard, "fused", iters=iters, warmup=6)
if rank == 0:
    print(json.dumps({"mode": "fused", "ms_per_iter": round(r["ms_per_iter"], 4), "checksum": r["x_checksum_first_16th"]}))
dist.destroy_process_group()
