"""Per-rank timing of the sharded power loop (development tool, run under torchrun): the shard's SpMV alone, the
python-driven NCCL halo loop, and the natively enqueued fused loop. Prints one line per rank."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from spmv_acc_b200 import sharded  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 384
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
loop, plan, csr = sharded.build_stencil3d_power_loop(N, "halo")
lo, hi = int(loop.bounds[rank]), int(loop.bounds[rank + 1])


def timed(fn, reps):
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(reps)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def alone(reps):
    for _ in range(reps):
        plan.execute(1.0, 0.0, loop.x, loop.x_next[lo:hi])


def tiles(reps):
    for _ in range(reps):
        for t0, t1 in loop.boundary:
            plan.execute_tiles(1.0, 0.0, loop.x, loop.x_next[lo:hi], t0, t1)
        for t0, t1 in loop.interior:
            plan.execute_tiles(1.0, 0.0, loop.x, loop.x_next[lo:hi], t0, t1)


alone(5)
t_alone = timed(alone, iters)
tiles(5)
t_tiles = timed(tiles, iters)
loop.run(4)
t_nccl = timed(loop.run, iters)
fused = sharded.FusedHaloLoop(loop, plan)
fused.run(4)
t_fused = timed(fused.run, iters)
info = plan.info()
print(f"[rank {rank}] rows={hi - lo} nnz={csr.nnz} tiles={info.ntiles} T={info.tile_nnz} boundary={loop.boundary} "
      f"alone={t_alone:.4f} ms  tile-ranges={t_tiles:.4f} ms  nccl-halo-loop={t_nccl:.4f} ms  fused-loop={t_fused:.4f} ms",
      flush=True)
fused.close()
dist.barrier()
dist.destroy_process_group()
