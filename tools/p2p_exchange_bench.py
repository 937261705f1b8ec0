"""Pure cost of the gradient exchange at N ranks (torchrun): PeerExchange.step back to back with nothing else on the GPU -- flag
barrier, fused reduce-scatter + Adam + all-gather kernel, flag barrier, gradient zeroing -- i.e. the exchange without any
skew between the ranks' training steps; and the bare flag barrier.  Scratch tool."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200._lib import call, stream
from ar_nerf_b200.sharding import PeerExchange
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 3072 + 2 * 5722520
px = PeerExchange(n, world, rank, dev)
p = torch.randn(n, device=dev); m = torch.zeros(px.owned, device=dev); v = torch.zeros(px.owned, device=dev)
px.grad.normal_()
step = [0]
def full():
    step[0] += 1
    px.step(p, m, v, (1e-2, 0.9, 0.999, 1e-15, step[0], 1.0), step[0], stream())
    px.wait_zeroed()
def barrier_only():
    step[0] += 1
    call("arn_p2p_barrier", px.F, px.f_ptr, world, rank, 0, step[0], stream())
def timeit(fn, iters=100):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
from ar_nerf_b200 import _lib
t_full = timeit(full)
t_bar = timeit(barrier_only, 300)
if rank == 0:
    print(f"world {world}: exchange step (barrier + kernel + barrier + zeroing, no skew) {t_full:.1f} us; bare barrier {t_bar:.1f} us; table {n * 4 / 1e6:.1f} MB; multicast {px.mc_grad is not None}")
if px.mc_grad is not None and os.environ.get("SWEEP"):
    for var in (1256, 1512, 2256, 2512, 4128, 4256, 8128):
        for g in (2, 4, 8):
            _lib.set_tunable("p2p_mc", var); px.grid_tail = g
            t = timeit(full, 60)
            if rank == 0:
                print(f"   p2p_mc {var} blocks/SM cap {g}: {t:.1f} us")
dist.destroy_process_group()
