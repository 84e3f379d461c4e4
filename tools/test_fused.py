"""Stage-by-stage check of the fused halo push on real peers (run under torchrun with 2+ ranks; development tool)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from spmv_acc_b200 import CsrDesc, PeerBuffer, SpmvPlan, stream_wait_flag, stream_write_flag, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def say(*a):
    print(f"[rank {rank}]", *a, flush=True)


N = 1 << 20
own = PeerBuffer.alloc(8 * N + 4 * world)
buf, flags = own.tensor("float64", N), own.tensor("int32", world, 8 * N)
buf.fill_(float(rank))
torch.cuda.synchronize()
objs = [None] * world
dist.all_gather_object(objs, (own.handle, own.nbytes))
peer = (rank + 1) % world
pb = PeerBuffer.open(*objs[peer])
say("peer buffer mapped at", hex(pb.address))

h = synth.stencil2d_numpy(64)
d = synth.to_device(h)
plan = SpmvPlan(CsrDesc(d.rows, d.cols, d.nnz, d.rowptr, d.col, d.val))
x = torch.ones(d.cols, dtype=torch.float64, device="cuda")
y = torch.zeros(d.rows, dtype=torch.float64, device="cuda")
plan.execute_push(1.0, 0.0, x, y, [(10, 50, pb.address + 8 * 1000)])
stream_write_flag(pb.address + 8 * N + 4 * rank, 7)
stream_wait_flag(flags.data_ptr() + 4 * ((rank - 1) % world), 7)
torch.cuda.synchronize()
say("rows 10..13 pushed by my left neighbour:", buf[1010:1014].tolist(), "expected", y[10:14].tolist(),
    "flags", flags.tolist())
assert torch.equal(buf[1010:1050], y[10:50])
dist.barrier()
pb.release()
dist.barrier()
dist.destroy_process_group()
say("done")
