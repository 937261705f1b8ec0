"""One 800x800 test frame (device-driven loop) and one SG shade / shadow-factor call: the command ncu wraps for the
test-render and SG kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.rendering import render
from ar_nerf_b200.sg_shadow import SGShadow
from ar_nerf_b200.workload import Workload
dev = torch.device("cuda:0")
w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model)
ro, rd = w.test_frame(800, 800); ro, rd = ro.to(dev), rd.to(dev)
for _ in range(2):
    r = render(model, ro, rd, test_time=True, T_threshold=1e-4)
g = torch.Generator(device="cpu").manual_seed(0)
sg = SGShadow.from_tensors(torch.randn(1, 128, 20, 20, 20, generator=g) * 0.15, torch.randn(128, 74, 148, generator=g) * 0.2,
                           torch.randn(1, 74, 148, generator=g) * 0.3, torch.rand(2048, 1024, generator=g), vol_range=2, device=dev)
axis = torch.nn.functional.normalize(torch.randn(32, 3, generator=g), dim=-1)
lSGs = torch.cat([axis, 10 ** (torch.rand(32, 1, generator=g) * 3.5 - 0.5), torch.rand(32, 3, generator=g) * 2 + 0.05], 1).to(dev)
n = 1920 * 1080
pts = (torch.rand(n, 3, generator=g) * 2 - 1).to(dev)
f = sg.calc_shadow_factor(0.3, pts, torch.zeros(3), lSGs)
m = 160000
nrm = torch.nn.functional.normalize(torch.randn(m, 3, generator=g), dim=-1).to(dev)
c = sg.shade(0.3, pts[:m].contiguous(), torch.zeros(3), lSGs, None, torch.rand(m, 3, generator=g).to(dev), torch.full((m, 1), 0.9, device=dev),
             torch.full((m, 1), 0.2, device=dev), nrm, -nrm, True)
torch.cuda.synchronize()
print("ok", int(r["total_samples"]), float(f.mean()), float(c.mean()))
