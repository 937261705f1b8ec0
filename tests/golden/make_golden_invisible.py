"""Golden vectors of NGP.mark_invisible_cells from the UNMODIFIED reference.

Run in the build container (needs /root/reference; CPU only):  python tests/golden/make_golden_invisible.py
models/networks.py cannot be imported here (it imports tinycudann), so the source of the ONE method is cut out of the file
with `ast` -- not a character of it is changed or copied into the repo -- compiled, and called on a stand-in `self` that has
exactly the attributes the method reads (cascades, grid_size, scale, density_grid, get_all_cells).  Inputs are regenerated
from seeds by invisible_inputs(); outputs go to tests/golden/mark_invisible_ref.npz (density as int8, the covered-camera
count as uint8: count_grid = count / N_cams)."""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/models/networks.py"


def invisible_inputs(case):
    """(grid_size, scale, cascades, K (3,3), poses (N,3,4), img_wh) of golden case 0 (bounded) / 1 (unbounded, 3 cascades)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from ar_nerf_b200.workload import intrinsics, look_at_poses
    if case == 0:
        return 32, 0.5, 1, intrinsics(800, 800), look_at_poses(20, 1.5, 7), (500, 420)  # the image covers part of the frustum only
    # cameras INSIDE the grid of the outer cascades: exercises the too-near rule and cells behind cameras
    return 32, 2.0, 3, intrinsics(200, 120), look_at_poses(12, 0.8, 11, upper_only=False), (200, 120)


def morton_cells(G):
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import oracle
    r = torch.arange(G, dtype=torch.int32)
    zz, yy, xx = torch.meshgrid(r, r, r, indexing='ij')
    coords = torch.stack([xx, yy, zz], -1).reshape(-1, 3).contiguous()
    return torch.from_numpy(oracle.morton3D(coords.numpy())).long(), coords


def reference_method():
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "NGP")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "mark_invisible_cells")
    mod = ast.Module(body=[fn], type_ignores=[])
    from einops import rearrange
    ns = {"torch": torch, "rearrange": rearrange, "NEAR_DISTANCE": 0.01}  # models/rendering.py:10
    exec(compile(mod, REF, "exec"), ns)
    return ns["mark_invisible_cells"]


def main():
    fn = reference_method()
    out = {}
    for case in (0, 1):
        G, scale, C, K, poses, wh = invisible_inputs(case)
        idx, coords = morton_cells(G)
        me = types.SimpleNamespace(cascades=C, grid_size=G, scale=scale, density_grid=torch.zeros(C, G ** 3),
                                   get_all_cells=lambda: [(idx, coords)] * C)
        fn(me, K, poses, wh, chunk=8192)
        out[f"density{case}"] = me.density_grid.numpy().astype(np.int8)
        cnt = me.count_grid.numpy() * len(poses)
        assert np.abs(cnt - np.round(cnt)).max() < 1e-3
        out[f"count{case}"] = np.round(cnt).astype(np.uint8)
        print(f"case {case}: {int((me.density_grid == 0).sum())} visible of {C * G ** 3} cells")
    np.savez_compressed(os.path.join(HERE, "mark_invisible_ref.npz"), **out)


if __name__ == "__main__":
    main()
