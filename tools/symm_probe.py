"""Does torch's symmetric memory give a multicast pointer on this box? (run under torchrun, development tool)"""
import json
import os

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
out = {"world": world}
try:
    out["has_multicast_support"] = bool(symm._SymmetricMemory.has_multicast_support("cuda", torch.cuda.current_device()))
except Exception as e:
    out["has_multicast_support_error"] = f"{type(e).__name__}: {e}"[:200]
try:
    t = symm.empty(1 << 20, dtype=torch.float64, device="cuda")
    h = symm.rendezvous(t, dist.group.WORLD)
    out.update({"rendezvous": True, "buffer_ptrs": [hex(p) for p in h.buffer_ptrs], "multicast_ptr": hex(h.multicast_ptr),
                "signal_pad_size": h.signal_pad_size, "buffer_size": h.buffer_size, "local_ptr": hex(t.data_ptr())})
    # write through the peer pointer of the next rank with a torch copy, read back locally
    t.fill_(float(rank))
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.float64)
    peer.fill_(100.0 + rank)
    h.barrier()
    torch.cuda.synchronize()
    out["first_after_peer_write"] = float(t[0].item())
except Exception as e:
    out["rendezvous_error"] = f"{type(e).__name__}: {e}"[:300]
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
