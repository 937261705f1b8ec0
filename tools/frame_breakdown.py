"""Where one 800x800 test frame spends its time: loop iterations, per-kernel device time, host share (scratch tool)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200 import _lib
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.rendering import render
from ar_nerf_b200.workload import Workload
dev = torch.device("cuda:0")
w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model)
ro, rd = w.test_frame(800, 800); ro, rd = ro.to(dev), rd.to(dev)
for thr, ms_ in ((1e-4, 1024), (1e-2, 100)):
    for _ in range(2):
        render(model, ro, rd, test_time=True, T_threshold=thr, max_samples=ms_)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = render(model, ro, rd, test_time=True, T_threshold=thr, max_samples=ms_)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
    n0 = _lib.launch_count()
    _lib.profile_enable(True)
    render(model, ro, rd, test_time=True, T_threshold=thr, max_samples=ms_)
    s = _lib.profile_report(); _lib.profile_enable(False)
    print(f"thr={thr} max_samples={ms_}: wall {wall:.2f} ms, launches {_lib.launch_count() - n0}, total samples {int(r['total_samples'])}")
    tot = 0
    for k, (n, ms) in sorted(s.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:32s} launches {n:4d}  ms {ms:7.3f}"); tot += ms
    print(f"   kernel total {tot:.3f} ms")
