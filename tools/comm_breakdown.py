"""Collective times at the sizes of the training step (torchrun): reduce-scatter of the fp32 table gradient, all-gather of
the fp16 table, the all-reduce they replace, the small MLP all-reduce."""
import os, sys
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 11448112
S = ((n + world - 1) // world + 7) // 8 * 8
P = S * world
g = torch.randn(P, device=dev); gs = torch.empty(S, device=dev)
p16 = torch.randn(P, device=dev).half()
small = torch.randn(7168, device=dev)
gb = g.bfloat16()
gsb = torch.empty(S, device=dev, dtype=torch.bfloat16)
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
res = {
    "reduce_scatter fp32 45.8MB": timeit(lambda: dist.reduce_scatter_tensor(gs, g)),
    "all_gather fp16 22.9MB": timeit(lambda: dist.all_gather_into_tensor(p16, p16[rank * S:(rank + 1) * S])),
    "all_reduce fp32 45.8MB": timeit(lambda: dist.all_reduce(g)),
    "reduce_scatter bf16 22.9MB": timeit(lambda: dist.reduce_scatter_tensor(gsb, gb)),
    "all_reduce fp32 28KB": timeit(lambda: dist.all_reduce(small)),
    "memset 45.8MB": timeit(lambda: g.zero_()),
}
if rank == 0:
    for k, v in res.items(): print(f"{k:32s} {v:8.1f} us")
dist.destroy_process_group()
