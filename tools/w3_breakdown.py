"""Per-kernel device time of the unbounded (W3: scale 16, 6 cascades, exp_step_factor 1/256) training step (scratch tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200 import _lib
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.trainer import NGPTrainer
from ar_nerf_b200.workload import Workload
dev = torch.device("cuda:0")
w = Workload("W3"); model = NGP(w.scale).to(dev); w.install(model)
tr = NGPTrainer(model, warmup_steps=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
B = [[t.to(dev) for t in w.train_batch(i)] for i in range(8)]
for i in range(40):
    tr.train_step(*B[i % 8][:3], next_rays=tuple(B[(i + 1) % 8][:2]))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(64):
    _, res = tr.train_step(*B[i % 8][:3], next_rays=tuple(B[(i + 1) % 8][:2]))
e1.record(); torch.cuda.synchronize()
print(f"W3 step: {e0.elapsed_time(e1) / 64:.3f} ms, samples/step {int(res['rm_samples'])}")
_lib.profile_enable(True)
for i in range(16):
    tr.train_step(*B[i % 8][:3], update_grid=False)
s = _lib.profile_report(); _lib.profile_enable(False)
for k, (n, ms) in sorted(s.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:30s} launches/step {n / 16:5.1f}  us/step {ms / 16 * 1e3:9.2f}")
e0.record()
for i in range(64):
    tr.train_step(*B[i % 8][:3], next_rays=tuple(B[(i + 1) % 8][:2]), update_grid=False)
e1.record(); torch.cuda.synchronize()
print(f"W3 step without refresh: {e0.elapsed_time(e1) / 64:.3f} ms")
_lib.profile_enable(True)
for k in range(4):
    for i in range(3):
        tr.train_step(*B[i % 8][:3], update_grid=False)
    model.update_density_grid(5.912, warmup=False)
s = _lib.profile_report(); _lib.profile_enable(False)
print("steady-state refresh, per refresh (us):", {k: round(ms / 4 * 1e3, 1) for k, (n, ms) in sorted(s.items(), key=lambda kv: -kv[1][1]) if 'train' not in k and 'composite' not in k and 'adam' not in k and 'bw' not in k})
