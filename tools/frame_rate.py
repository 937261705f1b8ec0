"""Device time of render(test_time=True) for the share of an 800x800 frame one rank of N renders (every N-th pixel), N = 1, 2,
4, 8: what the frame costs a rank before the gather.  Scratch tool (one GPU)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.rendering import render
from ar_nerf_b200.workload import Workload
dev = torch.device("cuda:0")
w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model)
ro, rd = w.test_frame(800, 800)
for world in (1, 2, 4, 8):
    o, d = ro[::world].contiguous().to(dev), rd[::world].contiguous().to(dev)
    ref = None
    variants = [{}, {"test_loop_launches": 4}, {"graph_test_loop": False}]
    variants += [{"samples_boost": k} for k in (2, 4, 8, 16, 32, 64)]
    for kw in variants:
        for _ in range(3):
            render(model, o, d, test_time=True, T_threshold=1e-4, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            r = render(model, o, d, test_time=True, T_threshold=1e-4, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        if ref is None:
            ref = r
        diff = max(float((ref[k] - r[k]).abs().max()) for k in ("rgb", "opacity"))
        from ar_nerf_b200 import rendering
        iters = [int(x['state_host'][0][5]) for x in rendering._TEST_WS.values()]
        print(f"[max |rgb, opacity - first line| {diff:.2e}; live iterations {iters}] ", end="")
        print(f"1/{world} of the frame ({o.shape[0]} rays), {kw or 'one loop'}: {ms:.3f} ms per frame = {1e3 / ms:.0f} frames/s, total samples {int(r['total_samples'])}")
