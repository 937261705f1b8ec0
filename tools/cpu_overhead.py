"""Host-side cost of NGPTrainer.train_step vs its device time (scratch tool)."""
import os, sys, time, cProfile, pstats
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.trainer import NGPTrainer
from ar_nerf_b200.workload import Workload
dev = torch.device("cuda:0")
w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model)
tr = NGPTrainer(model)
B = [[t.to(dev) for t in w.train_batch(i)] for i in range(8)]
def run(n, prefetch, upd=False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        b = B[i % 8]; nb = B[(i + 1) % 8]
        tr.train_step(b[0], b[1], b[2], update_grid=upd, next_rays=(nb[0], nb[1]) if prefetch else None)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3
for pf in (False, True):
    run(20, pf)
    print("prefetch", pf, "enqueue ms/step %.3f  total ms/step %.3f" % run(200, pf))
print("with grid update, prefetch: enqueue %.3f total %.3f" % run(192, True, True))
pr = cProfile.Profile(); pr.enable(); run(200, True); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
