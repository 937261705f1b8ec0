"""Per-entry-point device time of a few training steps / one test frame on the W1 workload (CUDA events around every
libarnerf.so call).  Scratch tool: numbers printed here are diagnostics, not bench values."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200 import _lib  # noqa: E402
from ar_nerf_b200.networks import NGP  # noqa: E402
from ar_nerf_b200.rendering import render  # noqa: E402
from ar_nerf_b200.trainer import NGPTrainer  # noqa: E402
from ar_nerf_b200.workload import Workload  # noqa: E402


def main():
    impl = sys.argv[1] if len(sys.argv) > 1 else ""
    dev = torch.device("cuda:0")
    w = Workload("W1")
    model = NGP(0.5).to(dev); model.field_impl = impl
    w.install(model)
    tr = NGPTrainer(model)
    batches = [[t.to(dev) for t in w.train_batch(i)] for i in range(8)]
    for i in range(3):
        tr.train_step(*batches[i][:3], noise=batches[i][3], update_grid=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, 8):
        _, res = tr.train_step(*batches[i][:3], noise=batches[i][3], update_grid=False)
    e1.record(); torch.cuda.synchronize()
    print(f"train step (no grid update, not instrumented): {e0.elapsed_time(e1) / 5:.3f} ms, samples/step {int(res['rm_samples'])}")
    _lib.profile_enable(True)
    for i in range(3, 8):
        tr.train_step(*batches[i][:3], noise=batches[i][3], update_grid=False)
    s = _lib.profile_report(); _lib.profile_enable(False)
    for k, (n, ms) in sorted(s.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:30s} launches/step {n / 5:5.1f}  us/step {ms / 5 * 1e3:9.2f}")
    print(f"  kernel total us/step {sum(ms for _, ms in s.values()) / 5 * 1e3:.1f}")
    _lib.TIMING = {}
    model.update_density_grid(5.912, warmup=True)
    s = _lib.timing_summary(); _lib.TIMING = None
    print("update_density_grid(warmup):", {k: round(v[1], 3) for k, v in s.items()})
    ro, rd = w.test_frame(800, 800)
    ro, rd = ro.to(dev), rd.to(dev)
    for thr, ms_ in ((1e-4, 1024), (1e-2, 100)):
        render(model, ro, rd, test_time=True, T_threshold=thr, max_samples=ms_)
        torch.cuda.synchronize(); e0.record()
        r = render(model, ro, rd, test_time=True, T_threshold=thr, max_samples=ms_)
        e1.record(); torch.cuda.synchronize()
        print(f"test frame 800x800 thr={thr} max_samples={ms_}: {e0.elapsed_time(e1):.2f} ms, total samples {int(r['total_samples'])}")


if __name__ == "__main__":
    main()
