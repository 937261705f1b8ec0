"""W4 (BASELINE.json configs[4]): one AR insertion frame at 1920x1080 on one GPU --
  (1) SG shading of the inserted object's G-buffer (400x400 bbox) under 32 SG lights with self shadow   [arn_sg_shade]
  (2) NeRF background render with the object's colours as IM_bkg and its depth as mesh_depth_map        [render(test_time=True, T=1e-2, 100 samples)]
  (3) the shadow the object casts on the scene, one factor per pixel of the frame                        [arn_sg_shadow_factor]
with synthetic stand-ins of the reference's git-ignored tables at the sizes insert/main.py:107 uses (f_h 2048x1024, PCA
volume 20^3 x 128 components, components 128 x 74 x 148).  Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200.networks import NGP  # noqa: E402
from ar_nerf_b200.workload import ARFrame, Workload  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model)
    fr = ARFrame(model, w, dev)

    def timed(fn, n=5):
        fn(); fn(); torch.cuda.synchronize()  # two untimed calls: workspaces and the allocator's large blocks exist
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out

    ms_frame, img = timed(fr.render)
    ms_shade, _ = timed(fr.shade, 20)
    pts_all = fr.ro + fr.rd * 1.0
    ms_factor, _ = timed(lambda: fr.sg.calc_shadow_factor(fr.model_r, pts_all, fr.model_pos, fr.lSGs), 20)
    line = {"workload": "W4 AR insertion frame 1920x1080: SG shading of a 400x400 object under 32 SG lights + NeRF background (T=1e-2, 100 samples) + SG shadow factor per pixel",
            "ms_per_frame": ms_frame, "frames_per_s": 1e3 / ms_frame, "sg_shade_ms": ms_shade, "sg_shade_pixels": fr.n_obj,
            "sg_shadow_factor_ms": ms_factor, "sg_shadow_factor_pixels": fr.H * fr.W, "finite": bool(torch.isfinite(img).all())}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
