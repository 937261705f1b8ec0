"""Per-window (16 steps, one occupancy refresh each) device time of the training loop -- looks for run-to-run variance."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.trainer import NGPTrainer
from ar_nerf_b200.workload import Workload
dev = torch.device("cuda:0")
w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model)
tr = NGPTrainer(model)
B = [[t.to(dev) for t in w.train_batch(i)] for i in range(16)]
pf = os.environ.get("PF", "1") == "1"
from ar_nerf_b200 import _lib
if os.environ.get("HBW"):
    _lib.set_tunable("hash_bw_mode", int(os.environ["HBW"]))
if os.environ.get("PARTS"):
    _lib.set_tunable("pipeline_parts", int(os.environ["PARTS"]))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
cpu = []
for win in range(10):
    ev[win].record(); t0 = time.perf_counter()
    for k in range(16):
        i = win * 16 + k
        b = B[i % 16]; nb = B[(i + 1) % 16]
        tr.train_step(b[0], b[1], b[2], next_rays=(nb[0], nb[1]) if pf else None)
    cpu.append((time.perf_counter() - t0) / 16 * 1e3)
ev[10].record(); torch.cuda.synchronize()
print("prefetch", pf)
print("gpu ms/step per window:", " ".join(f"{ev[i].elapsed_time(ev[i+1]) / 16:.3f}" for i in range(10)))
print("cpu enqueue ms/step   :", " ".join(f"{c:.3f}" for c in cpu))
print("memory allocated GB", torch.cuda.memory_allocated() / 1e9, "reserved", torch.cuda.memory_reserved() / 1e9, "num_alloc_retries", torch.cuda.memory_stats()["num_alloc_retries"], "segments", torch.cuda.memory_stats()["segment.all.allocated"])
