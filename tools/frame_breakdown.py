"""Per-kernel device time of ONE rank's share of an 800x800 test frame (every N-th pixel), device-driven loop without graphs so
that libarnerf's per-launch events can bracket every kernel.  Scratch tool (one GPU): python tools/frame_breakdown.py [world] [boost]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200 import _lib
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.rendering import render
from ar_nerf_b200.workload import Workload
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
boosts = [int(a) for a in sys.argv[2:]] or [1, 16]
dev = torch.device("cuda:0")
w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model)
ro, rd = w.test_frame(800, 800)
o, d = ro[::world].contiguous().to(dev), rd[::world].contiguous().to(dev)
for boost in boosts:
    kw = dict(test_time=True, T_threshold=1e-4, samples_boost=boost, graph_test_loop=False)
    for _ in range(3):
        render(model, o, d, **kw)
    torch.cuda.synchronize()
    _lib.profile_enable(True); _lib.profile_report()
    n = 10
    for _ in range(n):
        render(model, o, d, **kw)
    torch.cuda.synchronize()
    rep = _lib.profile_report(); _lib.profile_enable(False)
    tot = sum(ms for _, ms in rep.values())
    print(f"1/{world} of the frame ({o.shape[0]} rays), samples_boost {boost}: kernels sum {tot / n * 1e3:.1f} us per frame")
    for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:36s} n/frame={c / n:6.1f}  us/frame={ms / n * 1e3:8.1f}  avg={ms / c * 1e3:7.1f} us")
