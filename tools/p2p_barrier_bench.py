"""Pure latency of the peer-memory flag barrier (torchrun, >= 2 ranks): back-to-back barriers with nothing else on the
stream, and barriers behind a kernel that writes 45 MB (the state the training step's first barrier finds)."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200._lib import call, stream
from ar_nerf_b200.sharding import PeerExchange
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
px = PeerExchange(11448112, world, rank, dev)
step = [0]
def barrier(slot):
    step[0] += 1
    call("arn_p2p_barrier", px.F, px.f_ptr, world, rank, slot, step[0], stream())
def timeit(fn, iters=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
buf = torch.empty(11448112, device=dev)
t_b = timeit(lambda: barrier(0))
t_fill = timeit(lambda: buf.fill_(1.0))
t_fb = timeit(lambda: (buf.fill_(1.0), barrier(0)))
if rank == 0:
    print(f"world {world}: barrier alone {t_b:.1f} us; 45 MB fill {t_fill:.1f} us; fill + barrier {t_fb:.1f} us (barrier adds {t_fb - t_fill:.1f} us)")
dist.destroy_process_group()
