"""Small fused training steps + an occupancy refresh + a small test frame: the command compute-sanitizer wraps."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.rendering import render
from ar_nerf_b200.trainer import NGPTrainer
from ar_nerf_b200.workload import Workload
dev = torch.device("cuda:0")
kind = sys.argv[1] if len(sys.argv) > 1 else "W1"
w = Workload(kind, n_poses=8); model = NGP(w.scale).to(dev); w.install(model)
tr = NGPTrainer(model, update_interval=2, sample_capacity=600 * 1024)
B = [[t.to(dev) for t in w.train_batch(i, 600)] for i in range(4)]
for i in range(3):
    loss, res = tr.train_step(B[i][0], B[i][1], B[i][2], next_rays=(B[i + 1][0], B[i + 1][1]))
ro, rd = w.test_frame(48, 40)
r = render(model, ro.to(dev), rd.to(dev), test_time=True, T_threshold=1e-2, max_samples=64, exp_step_factor=w.exp_step_factor)
torch.cuda.synchronize()
print("ok", float(loss), int(res["rm_samples"]), int(r["total_samples"]))
