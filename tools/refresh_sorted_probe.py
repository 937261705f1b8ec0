"""Does the occupancy refresh's density evaluation get cheaper when its 1 M sampled cells are visited in curve order?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200 import _lib, vren  # noqa: E402
from ar_nerf_b200.networks import NGP  # noqa: E402
from ar_nerf_b200.workload import Workload  # noqa: E402

dev = torch.device("cuda:0")
w = Workload("W1")
model = NGP(0.5).to(dev)
w.install(model)
G = model.grid_size
M = G ** 3 // 4
for _ in range(20):
    model.update_density_grid(5.912)
coords1 = torch.randint(G, (M, 3), dtype=torch.int32, device=dev)
u = torch.randint(2 ** 31 - 1, (M,), device=dev)
rnd = torch.rand((2 * M, 3), device=dev)
idx, xyz = vren.grid_sample_cells(model.density_grid[0], 5.912, G, 0.5, coords1, u, rnd)
variants = {"draw order": xyz}
o = torch.argsort(idx); variants["sorted by cell"] = xyz[o].contiguous()
o = torch.argsort(idx >> 12); variants["bucketed by top 9 bits"] = xyz[o].contiguous()
o = torch.argsort(idx >> 9); variants["bucketed by top 12 bits"] = xyz[o].contiguous()
for name, x in variants.items():
    for _ in range(3):
        model.density(x)
    _lib.profile_enable(True)
    for _ in range(10):
        model.density(x)
    s = _lib.profile_report(); _lib.profile_enable(False)
    print(f"{name:28s}", {k: round(ms / 10 * 1e3, 1) for k, (n, ms) in s.items()})
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    torch.sort(idx)
e1.record(); torch.cuda.synchronize()
print("torch.sort of the 1 M int64 indices: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
