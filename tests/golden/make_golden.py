"""Generates tests/golden/vren_ref_*.npz by running the UNMODIFIED reference CUDA extension (oracle/_ref/vren.so,
built by oracle/build_ref.sh from /root/reference/models/csrc) on a B200:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy the npz files into tests/golden/

Inputs come from seeded ar_nerf_b200.workload generators; they are stored next to the outputs so the CPU oracle test
(tests/test_oracle_golden.py) needs nothing else."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refvren  # noqa: E402
from ar_nerf_b200.workload import Workload  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    vren = refvren.load()
    assert vren is not None, "oracle/_ref/vren*.so missing: run oracle/build_ref.sh"
    dev = "cuda"
    for kind, n_rays in (("W1", 384), ("W3", 256)):
        w = Workload(kind, n_poses=20)
        ro, rd, _, noise = w.train_batch(11, n_rays)
        # a few adversarial rays: axis-parallel, starting inside the box, missing the scene
        ro[:4] = torch.tensor([[0, 0, -2.0], [0.01, 0.02, 0.03], [3, 3, 3], [0.1, -2, 0.05]]) * (1 if kind == "W1" else 4)
        rd[:4] = torch.tensor([[0, 0, 1.0], [0.3, -0.2, 0.9], [1, 0, 0], [0, 1, 0]])
        g = {}
        ro_d, rd_d, noise_d = ro.to(dev), rd.to(dev), noise.to(dev)
        center = torch.zeros(1, 3, device=dev); half = torch.full((1, 3), w.scale, device=dev)
        cnt, hits_t, hidx = vren.ray_aabb_intersect(ro_d, rd_d, center, half, 1)
        g.update(rays_o=ro.numpy(), rays_d=rd.numpy(), noise=noise.numpy(), bitfield=w.bitfield.numpy(),
                 aabb_cnt=cnt.cpu().numpy(), aabb_hits_t=hits_t.cpu().numpy(), aabb_idx=hidx.cpu().numpy(),
                 cascades=w.cascades, scale=w.scale, esf=w.exp_step_factor)
        # multi-voxel intersection (8 voxels, max_hits 4)
        vc = (torch.rand(8, 3, generator=torch.Generator().manual_seed(3)) - 0.5) * w.scale
        vh = torch.full((8, 3), 0.2 * w.scale)
        cnt8, ht8, hi8 = vren.ray_aabb_intersect(ro_d, rd_d, vc.to(dev), vh.to(dev), 4)
        g.update(vox_centers=vc.numpy(), vox_half=vh.numpy(), vox_cnt=cnt8.cpu().numpy(), vox_hits_t=ht8.cpu().numpy())
        scnt, sht, shi = vren.ray_sphere_intersect(ro_d, rd_d, vc.to(dev), vh[:, 0].contiguous().to(dev), 4)
        g.update(sph_cnt=scnt.cpu().numpy(), sph_hits_t=sht.cpu().numpy())
        ht = hits_t[:, 0].clone()
        ht[(ht[:, 0] >= 0) & (ht[:, 0] < 0.01), 0] = 0.01
        bits = w.bitfield.to(dev)
        out = vren.raymarching_train(ro_d, rd_d, ht, bits, w.cascades, w.scale, w.exp_step_factor, noise_d, 128, 1024)
        n, xyzs, dirs, deltas, ts, total = refvren.canonical_train(out)
        g.update(train_n=n, train_xyzs=xyzs, train_deltas=deltas, train_ts=ts, train_total=total)
        # test-time marching, a few rounds of the rendering.py schedule
        hits = ht.clone(); alive = torch.arange(n_rays, device=dev)
        for i, S in enumerate((1, 2, 8, 64)):
            xyz_t, dir_t, dl_t, ts_t, neff = vren.raymarching_test(ro_d, rd_d, hits, alive, bits, w.cascades, w.scale,
                                                                   w.exp_step_factor, 128, 1024, S)
            g[f"test{i}_ts"] = ts_t.cpu().numpy(); g[f"test{i}_deltas"] = dl_t.cpu().numpy()
            g[f"test{i}_neff"] = neff.cpu().numpy(); g[f"test{i}_hits"] = hits.cpu().numpy(); g[f"test{i}_xyzs"] = xyz_t.cpu().numpy()
        # compositing fw/bw on the marched samples with seeded sigmas/rgbs (canonical layout)
        rays_a = torch.tensor(np.stack([np.arange(n_rays), np.concatenate([[0], np.cumsum(n)[:-1]]), n], 1), device=dev)
        gen = torch.Generator().manual_seed(5)
        N = total
        sig = (torch.rand(N, generator=gen) * 60).to(dev); rgbs = torch.rand(N, 3, generator=gen).to(dev)
        dl_d, ts_d = torch.tensor(deltas, device=dev), torch.tensor(ts, device=dev)
        for thr in (1e-4, 1e-2):
            tot, opa, dep, rgb, ws = vren.composite_train_fw(sig, rgbs, dl_d, ts_d, rays_a, thr)
            gO, gD = torch.randn(n_rays, generator=gen).to(dev), torch.randn(n_rays, generator=gen).to(dev)
            gC, gW = torch.randn(n_rays, 3, generator=gen).to(dev), torch.randn(N, generator=gen).to(dev)
            dsig, drgbs = vren.composite_train_bw(gO, gD, gC, gW, sig, rgbs, ws, dl_d, ts_d, rays_a, opa, dep, rgb, thr)
            k = f"comp{thr:g}_"
            g.update({k + "total": tot.cpu().numpy(), k + "opacity": opa.cpu().numpy(), k + "depth": dep.cpu().numpy(),
                      k + "rgb": rgb.cpu().numpy(), k + "ws": ws.cpu().numpy(), k + "gO": gO.cpu().numpy(), k + "gD": gD.cpu().numpy(),
                      k + "gC": gC.cpu().numpy(), k + "gW": gW.cpu().numpy(), k + "dsig": dsig.cpu().numpy(), k + "drgbs": drgbs.cpu().numpy()})
        g.update(comp_sigmas=sig.cpu().numpy(), comp_rgbs=rgbs.cpu().numpy())
        loss, wsi, wtsi = vren.distortion_loss_fw(ws, dl_d, ts_d, rays_a)
        gl = torch.randn(n_rays, generator=gen).to(dev)
        dws = vren.distortion_loss_bw(gl, wsi, wtsi, ws, dl_d, ts_d, rays_a)
        g.update(dist_loss=loss.cpu().numpy(), dist_wsi=wsi.cpu().numpy(), dist_wtsi=wtsi.cpu().numpy(), dist_gl=gl.cpu().numpy(),
                 dist_dws=dws.cpu().numpy(), dist_ws=ws.cpu().numpy())
        # grid utilities
        coords = torch.randint(0, 128, (4096, 3), dtype=torch.int32, generator=gen)
        g.update(morton_coords=coords.numpy(), morton_idx=vren.morton3D(coords.to(dev)).cpu().numpy())
        dens = torch.rand(2 * 128 ** 3 // 64, generator=gen) * 12
        bf = torch.zeros(dens.numel() // 8, dtype=torch.uint8, device=dev)
        vren.packbits(dens.to(dev), 5.912, bf)
        g.update(pack_density=dens.numpy().astype(np.float16).astype(np.float32), pack_thr=5.912)
        bf2 = torch.zeros_like(bf); vren.packbits(torch.tensor(g["pack_density"], device=dev), 5.912, bf2)
        g.update(pack_bits=bf2.cpu().numpy())
        path = os.path.join(out_dir, f"vren_ref_{kind}.npz")
        np.savez_compressed(path, **g)
        print("wrote", path, os.path.getsize(path), "bytes; total samples", total)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
