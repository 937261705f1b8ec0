"""W4 (BASELINE.json configs[4]): one AR insertion frame at 1920x1080 on one GPU --
  (1) SG shading of the inserted object's G-buffer (400x400 bbox) under 32 SG lights with self shadow   [arn_sg_shade]
  (2) NeRF background render with the object's colours as IM_bkg and its depth as mesh_depth_map        [render(test_time=True, T=1e-2, 100 samples)]
  (3) the shadow the object casts on the scene, one factor per pixel of the frame                        [arn_sg_shadow_factor]
with synthetic stand-ins of the reference's git-ignored tables at the sizes insert/main.py:107 uses (f_h 2048x1024, PCA
volume 20^3 x 128 components, components 128 x 74 x 148).  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200.networks import NGP  # noqa: E402
from ar_nerf_b200.rendering import render  # noqa: E402
from ar_nerf_b200.sg_shadow import SGShadow  # noqa: E402
from ar_nerf_b200.workload import Workload  # noqa: E402

H, W, BB = 1080, 1920, 400


def main():
    dev = torch.device("cuda:0")
    w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model)
    ro, rd = w.test_frame(H, W); ro, rd = ro.to(dev), rd.to(dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    sg = SGShadow.from_tensors(torch.randn(1, 128, 20, 20, 20, generator=g) * 0.15, torch.randn(128, 74, 148, generator=g) * 0.2,
                               torch.randn(1, 74, 148, generator=g) * 0.3, torch.rand(2048, 1024, generator=g), vol_range=2, device=dev)
    axis = torch.nn.functional.normalize(torch.randn(32, 3, generator=g), dim=-1)
    lSGs = torch.cat([axis, 10 ** (torch.rand(32, 1, generator=g) * 3.5 - 0.5), torch.rand(32, 3, generator=g) * 2 + 0.05], 1).to(dev)
    # G-buffer of a sphere-ish object in a BB x BB box at the centre of the frame
    ys, xs = torch.meshgrid(torch.arange(BB), torch.arange(BB), indexing="ij")
    rr = ((xs - BB / 2) ** 2 + (ys - BB / 2) ** 2).float().sqrt() / (BB / 2)
    inside = (rr < 1).flatten()
    n_obj = int(inside.sum())
    nz = (1 - rr.clamp(max=1) ** 2).sqrt()
    normal = torch.stack([(xs - BB / 2) / (BB / 2), -(ys - BB / 2) / (BB / 2), nz], -1).reshape(-1, 3)[inside].float().to(dev)
    box = torch.zeros(H, W, dtype=torch.bool); box[H // 2 - BB // 2:H // 2 + BB // 2, W // 2 - BB // 2:W // 2 + BB // 2] = inside.reshape(BB, BB)
    sel = box.flatten().to(dev)
    vdirs = torch.nn.functional.normalize(rd[sel], dim=-1)
    depth_obj = torch.full((n_obj,), 1.2, device=dev)
    pts_obj = ro[sel] + vdirs * depth_obj[:, None]
    albedo = torch.rand(n_obj, 3, generator=g).to(dev); metal = torch.full((n_obj, 1), 0.9, device=dev); rough = torch.full((n_obj, 1), 0.2, device=dev)
    model_pos = torch.tensor([0.0, 0.0, 0.0]); model_r = 0.3

    def frame():
        cols = sg.shade(model_r, pts_obj, model_pos, lSGs, None, albedo, metal, rough, normal, vdirs, True)       # main.py:559-576
        im_bkg = torch.zeros(H * W, 3, device=dev); im_bkg[sel] = cols
        mesh_depth = torch.zeros(H * W, device=dev); mesh_depth[sel] = depth_obj
        res = render(model, ro, rd, test_time=True, T_threshold=1e-2, max_samples=100, IM_bkg=im_bkg, mesh_depth_map=mesh_depth)  # main.py:646-650
        pts = ro + rd * res["depth"][:, None]                                                                          # main.py:493
        smap = sg.calc_shadow_factor(model_r, pts, model_pos, lSGs)                                                    # main.py:501
        return res["rgb"] * smap[:, None]

    def timed(fn, n=5):
        fn(); fn(); torch.cuda.synchronize()  # two untimed calls: workspaces and the allocator's large blocks exist
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out

    ms_frame, img = timed(frame)
    ms_shade, _ = timed(lambda: sg.shade(model_r, pts_obj, model_pos, lSGs, None, albedo, metal, rough, normal, vdirs, True), 20)
    pts_all = ro + rd * 1.0
    ms_factor, _ = timed(lambda: sg.calc_shadow_factor(model_r, pts_all, model_pos, lSGs), 20)
    line = {"workload": "W4 AR insertion frame 1920x1080: SG shading of a 400x400 object under 32 SG lights + NeRF background (T=1e-2, 100 samples) + SG shadow factor per pixel",
            "ms_per_frame": ms_frame, "frames_per_s": 1e3 / ms_frame, "sg_shade_ms": ms_shade, "sg_shade_pixels": n_obj,
            "sg_shadow_factor_ms": ms_factor, "sg_shadow_factor_pixels": H * W, "finite": bool(torch.isfinite(img).all())}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
