"""Probe: does torch's symmetric memory (cuMem + NVLS multicast) rendezvous work on this box?  2+ ranks, torchrun."""
import os, sys, json, traceback
import torch, torch.distributed as dist
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {"rank": rank}
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.uint8, device=dev)
    t.zero_()
    h = symm.rendezvous(t, dist.group.WORLD)
    out.update(multicast_ptr=int(h.multicast_ptr), buffer_ptrs=[int(p) for p in h.buffer_ptrs], local=int(t.data_ptr()),
               has_multicast=bool(getattr(h, "has_multicast_support", lambda *a: None) and h.multicast_ptr != 0),
               world=int(h.world_size), backend=str(symm.get_backend(dev)))
except Exception as e:
    out["error"] = "".join(traceback.format_exception_only(type(e), e))[-600:]
print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
