"""CPU oracle for the AR-NeRF rendering hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product (``ar_nerf_b200``) never does; it fails loudly when its CUDA library is missing.

``oracle_vren.c``   restates the reference's ``vren`` CUDA kernels (models/csrc) -- parity PINNED against the real
                    reference kernels (oracle/_ref, tests/golden).
``oracle_field.c``  restates the un-vendored tiny-cuda-nn pieces (hash grid, SH-4, 64-wide MLPs) -- parity UNPINNED.
``field_torch.py``  independent fp32 autograd restatement of the field, cross-checks the hand-derived backward.

numpy in / numpy out; function names follow the reference's pybind module (models/csrc/binding.cpp:234-250).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("oracle_vren.c", "oracle_field.c", "Makefile")):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_march_train_emit.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


f = C.c_float
i32 = C.c_int
i64 = C.c_int64


# ------------------------------------------------------------------ vren restatement
def ray_aabb_intersect(rays_o, rays_d, centers, half_sizes, max_hits):
    rays_o, rays_d, centers, half_sizes = map(_f32, (rays_o, rays_d, centers, half_sizes))
    R, V = len(rays_o), len(centers)
    cnt = np.zeros(R, np.int32); ht = np.zeros((R, max_hits, 2), np.float32); hi = np.zeros((R, max_hits), np.int64)
    lib().orc_ray_aabb_intersect(i32(R), _p(rays_o), _p(rays_d), i32(V), _p(centers), _p(half_sizes), i32(max_hits),
                                 _p(cnt), _p(ht), _p(hi))
    return cnt, ht, hi


def ray_sphere_intersect(rays_o, rays_d, centers, radii, max_hits):
    rays_o, rays_d, centers, radii = map(_f32, (rays_o, rays_d, centers, radii))
    R, V = len(rays_o), len(centers)
    cnt = np.zeros(R, np.int32); ht = np.zeros((R, max_hits, 2), np.float32); hi = np.zeros((R, max_hits), np.int64)
    lib().orc_ray_sphere_intersect(i32(R), _p(rays_o), _p(rays_d), i32(V), _p(centers), _p(radii), i32(max_hits),
                                   _p(cnt), _p(ht), _p(hi))
    return cnt, ht, hi


def morton3D(coords):
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    out = np.zeros(len(coords), np.int32)
    lib().orc_morton3d(i32(len(coords)), _p(coords), _p(out))
    return out


def morton3D_invert(indices):
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    out = np.zeros((len(indices), 3), np.int32)
    lib().orc_morton3d_invert(i32(len(indices)), _p(indices), _p(out))
    return out


def packbits(density_grid, threshold, density_bitfield):
    g = _f32(density_grid).reshape(-1)
    assert density_bitfield.dtype == np.uint8 and density_bitfield.flags.c_contiguous
    lib().orc_packbits(i32(density_bitfield.size), _p(g), f(threshold), _p(density_bitfield))


def mark_invisible_cells(coords, indices, grid_size, s, poses, K, img_wh, near=0.01, return_margin=False):
    """models/networks.py:209-250 for one cascade (half extent s): (density (G^3) f32 in {0,-1}, count (G^3) f32) written at
    `indices`.  float32 throughout, plain IEEE mul/add in a FIXED order (the reference's batched matmuls leave the order to
    cuBLAS / MKL): x_w = ((c * 1/(G-1)) * 2 - 1) * f32(s - s/G); x_c = ((R0 x + R1 y) + R2 z) + T; uvd = (K0 x_c + K1 y_c) + K2 z_c.
    return_margin: also the smallest relative distance of any decision (d>=0, d>=near, 0<=u<W, 0<=v<H) to its threshold per
    cell, evaluated in float64 -- cells with a tiny margin may legitimately differ under another summation order."""
    f32 = np.float32
    coords = np.ascontiguousarray(coords, np.int32); indices = _i64(indices)
    poses = _f32(poses); K = _f32(K)
    G = int(grid_size)
    smh = f32(float(s) - float(s) / G)
    xyz = ((coords.astype(f32) * f32(f32(1.0) / f32(G - 1))) * f32(2.0) - f32(1.0)) * smh          # (n,3)
    rot = np.transpose(poses[:, :3, :3], (0, 2, 1)).astype(f32)                                  # (N,3,3) world -> camera
    trans = np.zeros((len(poses), 3), f32)
    for i in range(3):  # -R^T t, torch bmm on tiny operands: computed in float64 and rounded once (the kernel receives this tensor)
        trans[:, i] = -(rot[:, i, :].astype(np.float64) * poses[:, :3, 3].astype(np.float64)).sum(1)
    n, N = len(coords), len(poses)
    covered = np.zeros(n, np.int32); too_near = np.zeros(n, bool)
    margin = np.full(n, np.inf)
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    W, H = f32(img_wh[0]), f32(img_wh[1])
    for c in range(N):
        R, T = rot[c], trans[c]
        pc = [((R[i, 0] * x + R[i, 1] * y) + R[i, 2] * z) + T[i] for i in range(3)]
        uvd = [(K[i, 0] * pc[0] + K[i, 1] * pc[1]) + K[i, 2] * pc[2] for i in range(3)]
        with np.errstate(divide="ignore", invalid="ignore"):
            u, v = uvd[0] / uvd[2], uvd[1] / uvd[2]
            d = uvd[2]
            in_image = (d >= 0) & (u >= 0) & (u < W) & (v >= 0) & (v < H)
            covered += (in_image & (d >= f32(near))).astype(np.int32)
            too_near |= in_image & (d < f32(near))
            if return_margin:
                d64, u64, v64 = d.astype(np.float64), u.astype(np.float64), v.astype(np.float64)
                m = np.minimum.reduce([np.abs(d64), np.abs(d64 - near), np.abs(u64) / float(W), np.abs(u64 - float(W)) / float(W),
                                       np.abs(v64) / float(H), np.abs(v64 - float(H)) / float(H)])
                margin = np.minimum(margin, np.nan_to_num(m, nan=0.0))
    count = covered.astype(f32) / f32(N)
    density = np.full(G ** 3, np.nan, f32); cnt = np.full(G ** 3, np.nan, f32)
    density[indices] = np.where((count > 0) & ~too_near, f32(0.0), f32(-1.0)); cnt[indices] = count
    if return_margin:
        mg = np.full(G ** 3, np.inf); mg[indices] = margin
        return density, cnt, mg
    return density, cnt


def world_to_camera(poses):
    """(N,12) f32: per camera the world-to-camera rotation row-major then the translation (what arn_mark_invisible_cells takes),
    with the translation computed as in mark_invisible_cells above."""
    poses = _f32(poses)
    rot = np.transpose(poses[:, :3, :3], (0, 2, 1)).astype(np.float32)
    trans = -(rot.astype(np.float64) * poses[:, None, :3, 3].astype(np.float64)).sum(2)
    return np.concatenate([rot.reshape(-1, 9), trans.astype(np.float32)], 1)


def raymarching_train(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise, grid_size,
                      max_samples):
    """Returns (rays_a, xyzs, dirs, deltas, ts, counter) already sliced to counter[0] samples, canonical ray order."""
    rays_o, rays_d, hits_t, noise = map(_f32, (rays_o, rays_d, hits_t, noise))
    bits = np.ascontiguousarray(density_bitfield, dtype=np.uint8)
    R = len(rays_o)
    n = np.zeros(R, np.int32)
    lib().orc_march_train_count(i32(R), _p(rays_o), _p(rays_d), _p(hits_t), _p(bits), i32(cascades), i32(grid_size),
                                f(scale), f(exp_step_factor), _p(noise), i32(max_samples), _p(n))
    tot = int(n.sum())
    rays_a = np.zeros((R, 3), np.int64)
    xyzs = np.zeros((tot, 3), np.float32); dirs = np.zeros((tot, 3), np.float32)
    deltas = np.zeros(tot, np.float32); ts = np.zeros(tot, np.float32)
    lib().orc_march_train_emit(i32(R), _p(rays_o), _p(rays_d), _p(hits_t), _p(bits), i32(cascades), i32(grid_size),
                               f(scale), f(exp_step_factor), _p(noise), i32(max_samples), _p(n), _p(rays_a),
                               _p(xyzs), _p(dirs), _p(deltas), _p(ts))
    return rays_a, xyzs, dirs, deltas, ts, np.array([tot, R], np.int32)


def raymarching_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale, exp_step_factor,
                     grid_size, max_samples, N_samples):
    """hits_t (R,2) float32 is updated IN PLACE (must be a C-contiguous float32 array)."""
    rays_o, rays_d = map(_f32, (rays_o, rays_d))
    assert hits_t.dtype == np.float32 and hits_t.flags.c_contiguous
    alive = _i64(alive_indices); bits = np.ascontiguousarray(density_bitfield, dtype=np.uint8)
    n = len(alive); S = N_samples
    xyzs = np.zeros((n, S, 3), np.float32); dirs = np.zeros((n, S, 3), np.float32)
    deltas = np.zeros((n, S), np.float32); ts = np.zeros((n, S), np.float32); neff = np.zeros(n, np.int32)
    lib().orc_march_test(i32(n), _p(rays_o), _p(rays_d), _p(hits_t), _p(alive), _p(bits), i32(cascades),
                         i32(grid_size), f(scale), f(exp_step_factor), i32(S), i32(max_samples),
                         _p(xyzs), _p(dirs), _p(deltas), _p(ts), _p(neff))
    return xyzs, dirs, deltas, ts, neff


def composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T_threshold):
    sigmas, rgbs, deltas, ts = map(_f32, (sigmas, rgbs, deltas, ts)); rays_a = _i64(rays_a)
    R, N = len(rays_a), len(sigmas)
    total = np.zeros(R, np.int64); opacity = np.zeros(R, np.float32); depth = np.zeros(R, np.float32)
    rgb = np.zeros((R, 3), np.float32); ws = np.zeros(N, np.float32)
    lib().orc_composite_train_fw(i32(R), _p(sigmas), _p(rgbs), _p(deltas), _p(ts), _p(rays_a), f(T_threshold),
                                 _p(total), _p(opacity), _p(depth), _p(rgb), _p(ws))
    return total, opacity, depth, rgb, ws


def composite_train_bw(dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, ws, deltas, ts, rays_a, opacity, depth,
                       rgb, T_threshold):
    (dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, ws, deltas, ts, opacity, depth, rgb) = map(
        _f32, (dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, ws, deltas, ts, opacity, depth, rgb))
    rays_a = _i64(rays_a)
    N = len(sigmas)
    dsig = np.zeros(N, np.float32); drgbs = np.zeros((N, 3), np.float32)
    lib().orc_composite_train_bw(i32(len(rays_a)), _p(dL_dopacity), _p(dL_ddepth), _p(dL_drgb), _p(dL_dws),
                                 _p(sigmas), _p(rgbs), _p(ws), _p(deltas), _p(ts), _p(rays_a), _p(opacity), _p(depth),
                                 _p(rgb), f(T_threshold), _p(dsig), _p(drgbs))
    return dsig, drgbs


def composite_test_fw(sigmas, rgbs, deltas, ts, hits_t, alive_indices, T_threshold, N_eff_samples, opacity, depth,
                      rgb):
    """In place on alive_indices (int64), opacity, depth, rgb (float32, C-contiguous)."""
    sigmas, rgbs, deltas, ts = map(_f32, (sigmas, rgbs, deltas, ts))
    for a, dt in ((alive_indices, np.int64), (opacity, np.float32), (depth, np.float32), (rgb, np.float32)):
        assert a.dtype == dt and a.flags.c_contiguous
    neff = np.ascontiguousarray(N_eff_samples, dtype=np.int32)
    n, S = sigmas.shape
    lib().orc_composite_test_fw(i32(n), i32(S), _p(sigmas), _p(rgbs), _p(deltas), _p(ts), _p(alive_indices),
                                f(T_threshold), _p(neff), _p(opacity), _p(depth), _p(rgb))


def distortion_loss_fw(ws, deltas, ts, rays_a):
    ws, deltas, ts = map(_f32, (ws, deltas, ts)); rays_a = _i64(rays_a)
    N, R = len(ws), len(rays_a)
    loss = np.zeros(R, np.float32); wsi = np.zeros(N, np.float32); wtsi = np.zeros(N, np.float32)
    lib().orc_distortion_fw(i32(R), i64(N), _p(ws), _p(deltas), _p(ts), _p(rays_a), _p(loss), _p(wsi), _p(wtsi))
    return loss, wsi, wtsi


def distortion_loss_bw(dL_dloss, ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts, rays_a):
    dL_dloss, wsi, wtsi, ws, deltas, ts = map(_f32, (dL_dloss, ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts))
    rays_a = _i64(rays_a)
    out = np.zeros(len(ws), np.float32)
    lib().orc_distortion_bw(i32(len(rays_a)), _p(dL_dloss), _p(wsi), _p(wtsi), _p(ws), _p(deltas), _p(ts), _p(rays_a),
                            _p(out))
    return out


# ------------------------------------------------------------------ tiny-cuda-nn restatement (parity unpinned)
class HashGeometry:
    """Level table, float32 host arithmetic."""

    def __init__(self, n_levels=16, base_resolution=16, per_level_scale=1.3195079, log2_hashmap_size=19):
        self.n_levels = n_levels
        self.scale = np.zeros(n_levels, np.float32); self.res = np.zeros(n_levels, np.uint32)
        self.size = np.zeros(n_levels, np.uint32); self.offset = np.zeros(n_levels + 1, np.uint32)
        lib().orc_hashgrid_geometry(i32(n_levels), i32(base_resolution), f(per_level_scale), i32(log2_hashmap_size),
                                    _p(self.scale), _p(self.res), _p(self.size), _p(self.offset))
        self.total = int(self.offset[-1])


def cast_f16(a):
    a = _f32(a); out = np.zeros(a.shape, np.float16)
    lib().orc_cast_f16(i64(a.size), _p(a), _p(out))
    return out


def hash_encode_fw(x01, geo, table16):
    x01 = _f32(x01); N = len(x01)
    feat = np.zeros((N, 32), np.float16)
    lib().orc_hash_encode_fw(i64(N), _p(x01), _p(geo.scale), _p(geo.res), _p(geo.size), _p(geo.offset),
                             _p(table16), _p(feat))
    return feat


def hash_encode_bw(x01, geo, table16, dfeat, want_table_grad=True, want_dx=False):
    x01 = _f32(x01); dfeat = _f32(dfeat); N = len(x01)
    tg = np.zeros((geo.total, 2), np.float64) if want_table_grad else None
    dx = np.zeros((N, 3), np.float32) if want_dx else None
    lib().orc_hash_encode_bw(i64(N), _p(x01), _p(geo.scale), _p(geo.res), _p(geo.size), _p(geo.offset),
                             _p(table16), _p(dfeat), _p(tg), _p(dx))
    return tg, dx


def sh4(dirs):
    dirs = _f32(dirs); out = np.zeros((len(dirs), 16), np.float16)
    lib().orc_sh4(i64(len(dirs)), _p(dirs), _p(out), i32(16))
    return out


def density_mlp_fw(feat16, Wd16):
    N = len(feat16)
    hid = np.zeros((N, 64), np.float16); h = np.zeros((N, 16), np.float32); sigma = np.zeros(N, np.float32)
    lib().orc_density_mlp_fw(i64(N), _p(feat16), _p(Wd16), _p(hid), _p(h), _p(sigma))
    return hid, h, sigma


def rgb_mlp_fw(sh16, h, Wc16, rgb_act=1):
    N = len(sh16)
    in32 = np.zeros((N, 32), np.float16); hid1 = np.zeros((N, 64), np.float16); hid2 = np.zeros((N, 64), np.float16)
    rgb = np.zeros((N, 3), np.float32)
    lib().orc_rgb_mlp_fw(i64(N), _p(sh16), _p(_f32(h)), _p(Wc16), i32(rgb_act), _p(in32), _p(hid1), _p(hid2), _p(rgb))
    return in32, hid1, hid2, rgb


def field_fw(x01, dirs, geo, params_xyz, params_rgb, rgb_act=1):
    """Whole field forward.  params_xyz fp32 flat [3072 MLP | table], params_rgb fp32 (7168).  Returns a ctx dict."""
    px = cast_f16(params_xyz); pc = cast_f16(params_rgb)
    Wd, table = px[:3072], px[3072:]
    feat = hash_encode_fw(x01, geo, table)
    hid, h, sigma = density_mlp_fw(feat, Wd)
    sh = sh4(dirs)
    in32, hid1, hid2, rgb = rgb_mlp_fw(sh, h, pc, rgb_act)
    return dict(x01=_f32(x01), feat=feat, hid=hid, h=h, sigma=sigma, in32=in32, hid1=hid1, hid2=hid2, rgb=rgb,
                Wd=Wd, Wc=pc, table=table, rgb_act=rgb_act)


def field_bw(ctx, geo, dL_dsigma, dL_drgb, loss_scale=128.0, want_dx=False):
    """Returns (grad_params_xyz fp32 flat, grad_params_rgb fp32, dL/dx01 or None)."""
    N = len(ctx["feat"])
    dWd = np.zeros(3072, np.float64); dWc = np.zeros(7168, np.float64); dfeat = np.zeros((N, 32), np.float32)
    lib().orc_field_mlp_bw(i64(N), _p(_f32(dL_dsigma)), _p(_f32(dL_drgb)), _p(ctx["rgb"]), _p(ctx["h"]),
                           _p(ctx["feat"]), _p(ctx["hid"]), _p(ctx["in32"]), _p(ctx["hid1"]), _p(ctx["hid2"]),
                           _p(ctx["Wd"]), _p(ctx["Wc"]), i32(ctx["rgb_act"]), f(loss_scale), _p(dWd), _p(dWc), _p(dfeat))
    tg, dx = hash_encode_bw(ctx["x01"], geo, ctx["table"], dfeat, True, want_dx)
    return np.concatenate([dWd, tg.reshape(-1)]), dWc, dx, dfeat


def field_bw_l1(ctx, geo, dL_dsigma, dL_drgb, loss_scale=128.0):
    """(l1_params_xyz, l1_params_rgb): for every parameter, the ALL-PATHS L1 of its gradient -- field_bw's walk with every
    weight replaced by |W| and every upstream gradient by |g| (orc_field_mlp_bw_l1), the hash-grid backward fed with the
    resulting L1 of dL/dfeat -- i.e. the sum of the absolute values of all products that reach the entry.  It is the scale
    against which a re-ordered or fp16-tie-perturbed sum is judged (tests/test_gpu_parity.py assert_sum): an entry is a sum
    of sums, and an inner sum that cancels carries its own uncertainty outward."""
    N = len(ctx["feat"])
    dWd = np.zeros(3072, np.float64); dWc = np.zeros(7168, np.float64); dfeat = np.zeros((N, 32), np.float32)
    lib().orc_field_mlp_bw_l1(i64(N), _p(_f32(dL_dsigma)), _p(_f32(dL_drgb)), _p(ctx["rgb"]), _p(ctx["h"]),
                              _p(ctx["feat"]), _p(ctx["hid"]), _p(ctx["in32"]), _p(ctx["hid1"]), _p(ctx["hid2"]),
                              _p(ctx["Wd"]), _p(ctx["Wc"]), i32(ctx["rgb_act"]), f(loss_scale), _p(dWd), _p(dWc), _p(dfeat))
    tg, _ = hash_encode_bw(ctx["x01"], geo, ctx["table"], dfeat, True, False)
    return np.concatenate([dWd, tg.reshape(-1)]), dWc
