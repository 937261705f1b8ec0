"""Does torch symmetric memory rendezvous here, and does it hand out a multicast (NVLS) address?  (torchrun, >= 2 ranks)"""
import os, sys
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    mc = getattr(h, "multicast_ptr", None)
    print(f"rank {rank}: rendezvous ok, buffer_ptrs {len(h.buffer_ptrs)} ptrs, multicast_ptr {hex(mc) if mc else mc}, signal_pad_ptrs {len(h.signal_pad_ptrs)}, "
          f"has_multicast_support {getattr(symm, 'has_multicast_support', lambda *a: 'n/a')('cuda', local) if hasattr(symm, 'has_multicast_support') else 'n/a'}", flush=True)
    t.fill_(rank + 1)
    dist.barrier(); torch.cuda.synchronize()
    peer = h.get_buffer((rank + 1) % world, (4,), torch.float32)
    print(f"rank {rank}: peer {(rank + 1) % world} buffer reads {peer.tolist()}", flush=True)
except Exception as e:
    print(f"rank {rank}: symmetric memory failed: {type(e).__name__}: {e}", flush=True)
dist.barrier()
dist.destroy_process_group()
