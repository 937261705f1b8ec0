"""Where update_density_grid spends its time (scratch tool)."""
import os, sys, time, cProfile, pstats
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.workload import Workload
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model)
for warm in (True, False):
    for _ in range(3): model.update_density_grid(5.912, warmup=warm)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): model.update_density_grid(5.912, warmup=warm)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"warmup={warm}: enqueue {(t1-t0)*100:.3f} ms  total {(t2-t0)*100:.3f} ms per update")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): model.update_density_grid(5.912, warmup=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
