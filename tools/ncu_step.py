"""A few fused training steps on W1 (no profiler hooks of our own): the command ncu wraps."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.trainer import NGPTrainer
from ar_nerf_b200.workload import Workload

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda:0")
w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model)
tr = NGPTrainer(model)
batches = [[t.to(dev) for t in w.train_batch(i)] for i in range(n_steps)]
for b in batches:
    loss, res = tr.train_step(*b[:3], noise=b[3], update_grid=False)
torch.cuda.synchronize()
print("ok", float(loss), int(res["rm_samples"]))
