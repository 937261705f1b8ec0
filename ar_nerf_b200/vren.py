"""`vren`-shaped module: the 12 functions of the reference's pybind11 extension (models/csrc/binding.cpp:234-250),
same names, argument order, return values and error behaviour, running on libarnerf.so (sm_100a) through ctypes.

Differences a caller can observe (all documented in DESIGN.md):
  * raymarching_train returns exactly-sized outputs in canonical ray order (rays_a[r] = (r, start, N)); the
    reference returns worst-case buffers in a scheduling-dependent order and the caller slices by counter[0].
  * kernels run on torch's current stream, not the legacy default stream.
"""
import torch

from . import _lib
from ._lib import F, I, L, call, check_tensor, ptr, stream

_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.float64: 2}
_T_SCRATCH = {}
_T_SCRATCH_LIMIT = 8 << 30  # bytes; larger requests fall back to the re-march emit


_COMPACT_SCAN = True  # tests flip it to cover the single-CTA scan over the rays_a rows


def _t_scratch(device, n):
    buf = _T_SCRATCH.get(device)
    if buf is None or buf.numel() < n:
        buf = torch.empty(n, dtype=torch.float32, device=device)
        _T_SCRATCH[device] = buf
    return buf


def ray_aabb_intersect(rays_o, rays_d, centers, half_sizes, max_hits):
    """binding.cpp:4-17 -> intersection.cu:59-100.  Returns (hit_cnt (R) i32, hits_t (R,max_hits,2), hits_voxel_idx (R,max_hits) i64)."""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (centers, "centers"), (half_sizes, "half_sizes")):
        check_tensor(t, n, torch.float32, 2, 3)
    R, V = rays_o.shape[0], centers.shape[0]
    dev = rays_o.device
    hit_cnt = torch.empty(R, dtype=torch.int32, device=dev)
    hits_t = torch.empty(R, max_hits, 2, dtype=torch.float32, device=dev)
    hits_idx = torch.empty(R, max_hits, dtype=torch.int64, device=dev)
    call("arn_ray_aabb_intersect", ptr(rays_o), ptr(rays_d), R, ptr(centers), ptr(half_sizes), V, int(max_hits),
         ptr(hit_cnt), ptr(hits_t), ptr(hits_idx), stream())
    return [hit_cnt, hits_t, hits_idx]


def ray_sphere_intersect(rays_o, rays_d, centers, radii, max_hits):
    """binding.cpp:19-32 -> intersection.cu:156-197."""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (centers, "centers")):
        check_tensor(t, n, torch.float32, 2, 3)
    check_tensor(radii, "radii", torch.float32, 1)
    R, V = rays_o.shape[0], centers.shape[0]
    dev = rays_o.device
    hit_cnt = torch.empty(R, dtype=torch.int32, device=dev)
    hits_t = torch.empty(R, max_hits, 2, dtype=torch.float32, device=dev)
    hits_idx = torch.empty(R, max_hits, dtype=torch.int64, device=dev)
    call("arn_ray_sphere_intersect", ptr(rays_o), ptr(rays_d), R, ptr(centers), ptr(radii), V, int(max_hits),
         ptr(hit_cnt), ptr(hits_t), ptr(hits_idx), stream())
    return [hit_cnt, hits_t, hits_idx]


def ray_aabb_near(rays_o, rays_d, center_host, half_size_host, near):
    """Fused rendering.py:29-31 (single box, max_hits=1, near clamp).  center/half_size are 3-float host sequences."""
    check_tensor(rays_o, "rays_o", torch.float32, 2, 3); check_tensor(rays_d, "rays_d", torch.float32, 2, 3)
    R = rays_o.shape[0]
    hits_t = torch.empty(R, 1, 2, dtype=torch.float32, device=rays_o.device)
    c = (F * 3)(*[float(v) for v in center_host]); h = (F * 3)(*[float(v) for v in half_size_host])
    call("arn_ray_aabb_near", ptr(rays_o), ptr(rays_d), R, c, h, float(near), ptr(hits_t), stream())
    return hits_t


def morton3D(coords):
    """binding.cpp:46-50 -> raymarching.cu:72-88."""
    check_tensor(coords, "coords", torch.int32, 2, 3)
    out = torch.empty(coords.shape[0], dtype=torch.int32, device=coords.device)
    call("arn_morton3d", ptr(coords), coords.shape[0], ptr(out), stream())
    return out


def morton3D_invert(indices):
    """binding.cpp:53-57 -> raymarching.cu:103-119."""
    check_tensor(indices, "indices", torch.int32, 1)
    out = torch.empty(indices.shape[0], 3, dtype=torch.int32, device=indices.device)
    call("arn_morton3d_invert", ptr(indices), indices.shape[0], ptr(out), stream())
    return out


def packbits(density_grid, density_threshold, density_bitfield):
    """binding.cpp:34-43 -> raymarching.cu:143-162.  In place on density_bitfield (uint8, C*G^3/8)."""
    check_tensor(density_grid, "density_grid"); check_tensor(density_bitfield, "density_bitfield", torch.uint8)
    if density_grid.dtype not in _DTYPE_CODE:
        raise RuntimeError(f'"packbits_cu" not implemented for \'{density_grid.dtype}\'')
    n_bytes = density_bitfield.numel()
    if density_grid.numel() < 8 * n_bytes:
        raise RuntimeError("density_grid must hold 8 cells per bitfield byte")
    call("arn_packbits", ptr(density_grid), _DTYPE_CODE[density_grid.dtype], float(density_threshold),
         ptr(density_bitfield), n_bytes, stream())


def gather_rays(poses, img_idxs, pix_idxs, directions=None, K=None, width=None):
    """rays_o, rays_d (n,3) of a sampled batch, on the device (train.py:121-126 + ray_utils.get_rays in one kernel).
    poses (n_images,3,4) f32; img_idxs (n) i64 tensor or a python int ('same_image'); pix_idxs (n) i64; directions (H*W,3) f32
    or K (3,3) + width to compute pixel-centre directions on the fly."""
    import ctypes as C
    check_tensor(poses, "poses", torch.float32, 3, 4); check_tensor(pix_idxs, "pix_idxs", torch.int64, 1)
    n, dev = pix_idxs.shape[0], poses.device
    single = 0
    if isinstance(img_idxs, torch.Tensor):
        check_tensor(img_idxs, "img_idxs", torch.int64, 1)
        if img_idxs.shape[0] != n:
            raise RuntimeError("img_idxs and pix_idxs must have the same length")
    else:
        single, img_idxs = int(img_idxs), None
    k_host = None
    if directions is not None:
        check_tensor(directions, "directions", torch.float32, 2, 3)
    else:
        if K is None or width is None:
            raise RuntimeError("gather_rays needs either directions or K and width")
        k_host = (C.c_float * 9)(*[float(v) for v in torch.as_tensor(K).reshape(-1).tolist()])
    rays_o = torch.empty(n, 3, dtype=torch.float32, device=dev); rays_d = torch.empty(n, 3, dtype=torch.float32, device=dev)
    call("arn_gather_rays", ptr(directions), k_host, int(width or 0), ptr(poses), ptr(img_idxs), single, ptr(pix_idxs), n,
         ptr(rays_o), ptr(rays_d), stream())
    return rays_o, rays_d


def gather_batch(poses, img_idxs, pix_idxs, images, directions=None, K=None, width=None, out=None):
    """gather_rays + the batch's pixels images[img_idxs, pix_idxs] (datasets/base.py:32) in one kernel.  images (n_images, H*W, C)
    f32 on the device.  out = (rays_o, rays_d, pixels) preallocated tensors to fill (the training loop's batch slots)."""
    import ctypes as C
    check_tensor(poses, "poses", torch.float32, 3, 4); check_tensor(pix_idxs, "pix_idxs", torch.int64, 1)
    check_tensor(images, "images", torch.float32, 3)
    n, dev = pix_idxs.shape[0], poses.device
    single = 0
    if isinstance(img_idxs, torch.Tensor):
        check_tensor(img_idxs, "img_idxs", torch.int64, 1)
        if img_idxs.shape[0] != n:
            raise RuntimeError("img_idxs and pix_idxs must have the same length")
    else:
        single, img_idxs = int(img_idxs), None
    k_host = None
    if directions is not None:
        check_tensor(directions, "directions", torch.float32, 2, 3)
    else:
        if K is None or width is None:
            raise RuntimeError("gather_batch needs either directions or K and width")
        k_host = (C.c_float * 9)(*[float(v) for v in torch.as_tensor(K).reshape(-1).tolist()])
    ch = images.shape[2]
    if out is None:
        out = (torch.empty(n, 3, dtype=torch.float32, device=dev), torch.empty(n, 3, dtype=torch.float32, device=dev),
               torch.empty(n, ch, dtype=torch.float32, device=dev))
    rays_o, rays_d, pixels = out
    for t, nm, last in ((rays_o, "rays_o", 3), (rays_d, "rays_d", 3), (pixels, "pixels", ch)):
        check_tensor(t, nm, torch.float32, 2, last)
        if t.shape[0] != n:
            raise RuntimeError(f"{nm} must have {n} rows")
    call("arn_gather_batch", ptr(directions), k_host, int(width or 0), ptr(poses), ptr(img_idxs), single, ptr(pix_idxs), n,
         ptr(images), images.shape[1], ch, ptr(rays_o), ptr(rays_d), ptr(pixels), stream())
    return rays_o, rays_d, pixels


def grid_cell_positions(coords, rnd, grid_size, s):
    """networks.py:263-267 in one kernel: (coords/(G-1)*2-1)*(s - s/G) + (rnd*2-1)*(s/G); coords (M,3) i32, rnd (M,3) f32."""
    check_tensor(coords, "coords", torch.int32, 2, 3); check_tensor(rnd, "rnd", torch.float32, 2, 3)
    if rnd.shape != coords.shape:
        raise RuntimeError("rnd must have the shape of coords")
    out = torch.empty(coords.shape, dtype=torch.float32, device=coords.device)
    call("arn_grid_cell_positions", ptr(coords), ptr(rnd), coords.shape[0], int(grid_size), float(s), ptr(out), stream())
    return out


_SAMPLE_SCRATCH = {}


def grid_sample_scratch_ints(grid_size, M, sort=False):
    """ARN_GRID_SAMPLE_SCRATCH_INTS / ARN_GRID_SAMPLE_SORTED_SCRATCH_INTS (include/arnerf.h)."""
    chunks = (grid_size ** 3 + 1023) // 1024
    return chunks * 36 + 12 + grid_size ** 3 + 2 * M if sort else chunks * 34 + 4


def grid_sample_cells(density_grid_c, density_threshold, grid_size, s, coords1, u, rnd, sort=False, out=None, scratch=None):
    """Steady-state cell selection of ONE cascade (networks.py:181-207 + :263-267) without torch glue: coords1 (M,3) i32 and
    u (M,) i64 are the caller's two randint draws, rnd (2M,3) its rand draw.  Returns (indices (2M,) i64, xyzs_w (2M,3)) in
    draw order (uniform half first) or, sort=True, along the morton curve (arn_grid_sample_cells_sorted).  out = (indices,
    xyzs) to write into; scratch = the caller's own int32 workspace (grid_sample_scratch_ints) instead of the shared one."""
    check_tensor(density_grid_c, "density_grid", torch.float32, 1)
    check_tensor(coords1, "coords1", torch.int32, 2, 3); check_tensor(u, "u", torch.int64, 1); check_tensor(rnd, "rnd", torch.float32, 2, 3)
    M = coords1.shape[0]
    if u.shape[0] != M or rnd.shape[0] != 2 * M or density_grid_c.numel() != grid_size ** 3:
        raise RuntimeError("grid_sample_cells: inconsistent sizes")
    dev = density_grid_c.device
    need = grid_sample_scratch_ints(grid_size, M, sort)
    if scratch is not None:  # the caller's own (calls that may overlap on different streams must not share one)
        check_tensor(scratch, "scratch", torch.int32, 1)
        if scratch.numel() < need:
            raise RuntimeError("grid_sample_cells: scratch too small")
    else:
        key = (dev.index, need, bool(sort))
        scratch = _SAMPLE_SCRATCH.get(key)
        if scratch is None:
            for k in [k for k in _SAMPLE_SCRATCH if k[2] == bool(sort)]:
                del _SAMPLE_SCRATCH[k]
            scratch = _SAMPLE_SCRATCH[key] = torch.empty(need, dtype=torch.int32, device=dev)
    if out is None:
        indices = torch.empty(2 * M, dtype=torch.int64, device=dev)
        xyzs = torch.empty(2 * M, 3, dtype=torch.float32, device=dev)
    else:
        indices, xyzs = out
        check_tensor(indices, "indices", torch.int64, 1); check_tensor(xyzs, "xyzs", torch.float32, 2, 3)
        if indices.shape[0] != 2 * M or xyzs.shape[0] != 2 * M:
            raise RuntimeError("grid_sample_cells: out tensors must hold 2M cells")
    call("arn_grid_sample_cells_sorted" if sort else "arn_grid_sample_cells", ptr(density_grid_c), float(density_threshold), int(grid_size), float(s),
         ptr(coords1), ptr(u), M, ptr(rnd), ptr(scratch), ptr(indices), ptr(xyzs), stream())
    return indices, xyzs


def grid_scatter(dst, indices, src):
    """dst[indices] = src for 1-D f32 dst / src and i64 indices (networks.py:268), one native launch."""
    check_tensor(dst, "dst", torch.float32, 1); check_tensor(indices, "indices", torch.int64, 1); check_tensor(src, "src", torch.float32, 1)
    if indices.shape[0] != src.shape[0]:
        raise RuntimeError("grid_scatter: indices and src must have the same length")
    call("arn_grid_scatter", ptr(dst), dst.numel(), ptr(indices), ptr(src), src.shape[0], stream())


_GRID_SCRATCH = {}


def density_grid_update(density_grid, density_tmp, decay_cells, decay, density_threshold, density_bitfield):
    """networks.py:273-281 without the host round trip: EMA/max refresh in place, threshold = min(mean(grid[grid>0]),
    density_threshold) on the device, packbits.  Returns the 1-element device tensor holding the threshold used."""
    check_tensor(density_grid, "density_grid", torch.float32); check_tensor(density_tmp, "density_tmp", torch.float32)
    check_tensor(density_bitfield, "density_bitfield", torch.uint8)
    n = density_grid.numel()
    if density_tmp.numel() != n or density_bitfield.numel() * 8 != n:
        raise RuntimeError("density_tmp / density_bitfield do not match density_grid")
    if decay_cells is not None:
        check_tensor(decay_cells, "decay_cells", torch.float32)
        if decay_cells.numel() != n:
            raise RuntimeError("decay_cells does not match density_grid")
    key = (density_grid.device.type, density_grid.device.index)
    if key not in _GRID_SCRATCH:
        _GRID_SCRATCH[key] = torch.empty(2048 * 16 + 16, dtype=torch.uint8, device=density_grid.device)
    scratch = _GRID_SCRATCH[key]
    call("arn_density_grid_update", ptr(density_grid), ptr(density_tmp), ptr(decay_cells), float(decay) if decay_cells is None else 0.0,
         float(density_threshold), n, ptr(density_bitfield), ptr(scratch), stream())
    return scratch[2048 * 16:2048 * 16 + 4].view(torch.float32)


def mark_invisible_cells(coords, indices, grid_size, s, w2c, K, img_wh, near, density_grid_c, count_grid_c):
    """networks.py:209-250 for one cascade, every camera in one launch.  coords (n,3) i32, indices (n) i64, w2c (n_cams,12) f32
    (rotation row-major | translation), K (3,3); density_grid_c / count_grid_c (G^3) f32 are written at `indices`."""
    import ctypes as C
    check_tensor(coords, "coords", torch.int32, 2, 3); check_tensor(indices, "indices", torch.int64, 1)
    check_tensor(w2c, "w2c", torch.float32, 2, 12)
    check_tensor(density_grid_c, "density_grid", torch.float32, 1); check_tensor(count_grid_c, "count_grid", torch.float32, 1)
    if indices.shape[0] != coords.shape[0]:
        raise RuntimeError("indices and coords must have the same length")
    k_host = (C.c_float * 9)(*[float(v) for v in torch.as_tensor(K, dtype=torch.float32).reshape(-1).tolist()])
    call("arn_mark_invisible_cells", ptr(coords), ptr(indices), coords.shape[0], int(grid_size), float(s), ptr(w2c), w2c.shape[0], k_host,
         float(img_wh[0]), float(img_wh[1]), float(near), ptr(density_grid_c), ptr(count_grid_c), stream())


def raymarching_train(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise, grid_size,
                      max_samples):
    """binding.cpp:60-81 -> raymarching.cu:283-332.  Returns [rays_a, xyzs, dirs, deltas, ts, counter]."""
    check_tensor(rays_o, "rays_o", torch.float32, 2, 3); check_tensor(rays_d, "rays_d", torch.float32, 2, 3)
    check_tensor(hits_t, "hits_t", torch.float32, 2, 2); check_tensor(density_bitfield, "density_bitfield", torch.uint8)
    check_tensor(noise, "noise", torch.float32, 1)
    R = rays_o.shape[0]
    dev = rays_o.device
    if density_bitfield.numel() < cascades * grid_size ** 3 // 8:
        raise RuntimeError("density_bitfield smaller than cascades*grid_size^3/8")
    rays_a = torch.empty(R, 3, dtype=torch.int64, device=dev)
    counter = torch.empty(2, dtype=torch.int32, device=dev)
    scratch_n = R * max_samples
    t_scratch = _t_scratch(dev, scratch_n) if 0 < scratch_n * 4 <= _T_SCRATCH_LIMIT else None
    cfg = (ptr(density_bitfield), int(cascades), int(grid_size), float(scale), float(exp_step_factor), ptr(noise),
           int(max_samples))
    count_scratch = torch.empty(R, dtype=torch.int32, device=dev) if _COMPACT_SCAN else None
    call("arn_march_train_count_ex", ptr(rays_o), ptr(rays_d), ptr(hits_t), R, *cfg, ptr(rays_a), ptr(counter),
         ptr(t_scratch), ptr(count_scratch), stream())
    total = int(counter[0].item()) if R > 0 else 0  # the one host sync the reference has too (custom_functions.py:91-96)
    xyzs = torch.empty(total, 3, dtype=torch.float32, device=dev)
    dirs = torch.empty(total, 3, dtype=torch.float32, device=dev)
    deltas = torch.empty(total, dtype=torch.float32, device=dev)
    ts = torch.empty(total, dtype=torch.float32, device=dev)
    call("arn_march_train_emit_ex", ptr(rays_o), ptr(rays_d), ptr(hits_t), R, *cfg, ptr(rays_a), ptr(t_scratch),
         ptr(xyzs), ptr(dirs), ptr(deltas), ptr(ts), total, stream())
    return [rays_a, xyzs, dirs, deltas, ts, counter]


def raymarching_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale, exp_step_factor,
                     grid_size, max_samples, N_samples):
    """binding.cpp:84-106 -> raymarching.cu:407-454.  hits_t (R,2) updated in place.  Returns [xyzs, dirs, deltas, ts, N_eff]."""
    check_tensor(rays_o, "rays_o", torch.float32, 2, 3); check_tensor(rays_d, "rays_d", torch.float32, 2, 3)
    check_tensor(hits_t, "hits_t", torch.float32, 2, 2); check_tensor(alive_indices, "alive_indices", torch.int64, 1)
    check_tensor(density_bitfield, "density_bitfield", torch.uint8)
    n, S, dev = alive_indices.shape[0], int(N_samples), rays_o.device
    xyzs = torch.empty(n, S, 3, dtype=torch.float32, device=dev)
    dirs = torch.empty(n, S, 3, dtype=torch.float32, device=dev)
    deltas = torch.empty(n, S, dtype=torch.float32, device=dev)
    ts = torch.empty(n, S, dtype=torch.float32, device=dev)
    n_eff = torch.empty(n, dtype=torch.int32, device=dev)
    call("arn_march_test", ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(alive_indices), n, ptr(density_bitfield),
         int(cascades), int(grid_size), float(scale), float(exp_step_factor), S, int(max_samples),
         ptr(xyzs), ptr(dirs), ptr(deltas), ptr(ts), ptr(n_eff), stream())
    return [xyzs, dirs, deltas, ts, n_eff]


def composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T_threshold):
    """binding.cpp:109-126 -> volumerendering.cu:47-83.  Returns [total_samples (R) i64, opacity, depth, rgb, ws]."""
    check_tensor(sigmas, "sigmas", torch.float32, 1); check_tensor(rgbs, "rgbs", torch.float32, 2, 3)
    check_tensor(deltas, "deltas", torch.float32, 1); check_tensor(ts, "ts", torch.float32, 1)
    check_tensor(rays_a, "rays_a", torch.int64, 2, 3)
    R, N, dev = rays_a.shape[0], sigmas.shape[0], sigmas.device
    total = torch.empty(R, dtype=torch.int64, device=dev)
    opacity = torch.empty(R, dtype=torch.float32, device=dev)
    depth = torch.empty(R, dtype=torch.float32, device=dev)
    rgb = torch.empty(R, 3, dtype=torch.float32, device=dev)
    ws = torch.empty(N, dtype=torch.float32, device=dev)
    call("arn_composite_train_fw", ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(ts), ptr(rays_a), R, N, float(T_threshold),
         ptr(total), ptr(opacity), ptr(depth), ptr(rgb), ptr(ws), stream())
    return [total, opacity, depth, rgb, ws]


def composite_train_bw(dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, ws, deltas, ts, rays_a, opacity, depth,
                       rgb, T_threshold):
    """binding.cpp:129-163 -> volumerendering.cu:153-201.  dL_dws may be None.  Returns [dL_dsigmas, dL_drgbs]."""
    for t, n in ((dL_dopacity, "dL_dopacity"), (dL_ddepth, "dL_ddepth"), (sigmas, "sigmas"), (ws, "ws"),
                 (deltas, "deltas"), (ts, "ts"), (opacity, "opacity"), (depth, "depth")):
        check_tensor(t, n, torch.float32, 1)
    check_tensor(dL_drgb, "dL_drgb", torch.float32, 2, 3); check_tensor(rgbs, "rgbs", torch.float32, 2, 3)
    check_tensor(rgb, "rgb", torch.float32, 2, 3); check_tensor(rays_a, "rays_a", torch.int64, 2, 3)
    if dL_dws is not None:
        check_tensor(dL_dws, "dL_dws", torch.float32, 1)
    R, N, dev = rays_a.shape[0], sigmas.shape[0], sigmas.device
    dsig = torch.empty(N, dtype=torch.float32, device=dev)
    drgbs = torch.empty(N, 3, dtype=torch.float32, device=dev)
    call("arn_composite_train_bw", ptr(dL_dopacity), ptr(dL_ddepth), ptr(dL_drgb), ptr(dL_dws), ptr(sigmas), ptr(rgbs),
         ptr(ws), ptr(deltas), ptr(ts), ptr(rays_a), ptr(opacity), ptr(depth), ptr(rgb), R, N, float(T_threshold),
         ptr(dsig), ptr(drgbs), stream())
    return [dsig, drgbs]


def composite_test_fw(sigmas, rgbs, deltas, ts, hits_t, alive_indices, T_threshold, N_eff_samples, opacity, depth, rgb):
    """binding.cpp:166-194 -> volumerendering.cu:251-284.  In place on alive_indices / opacity / depth / rgb."""
    check_tensor(sigmas, "sigmas", torch.float32, 2); check_tensor(rgbs, "rgbs", torch.float32, 3, 3)
    check_tensor(deltas, "deltas", torch.float32, 2); check_tensor(ts, "ts", torch.float32, 2)
    check_tensor(hits_t, "hits_t")  # accepted and unused, as in the reference kernel
    check_tensor(alive_indices, "alive_indices", torch.int64, 1); check_tensor(N_eff_samples, "N_eff_samples", torch.int32, 1)
    check_tensor(opacity, "opacity", torch.float32, 1); check_tensor(depth, "depth", torch.float32, 1)
    check_tensor(rgb, "rgb", torch.float32, 2, 3)
    n, S = sigmas.shape
    call("arn_composite_test_fw", ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(ts), ptr(alive_indices), n, S,
         float(T_threshold), ptr(N_eff_samples), ptr(opacity), ptr(depth), ptr(rgb), stream())


def distortion_loss_fw(ws, deltas, ts, rays_a):
    """binding.cpp:197-209 -> losses.cu:62-107.  Returns [loss (R), ws_inclusive_scan (N), wts_inclusive_scan (N)]."""
    for t, n in ((ws, "ws"), (deltas, "deltas"), (ts, "ts")):
        check_tensor(t, n, torch.float32, 1)
    check_tensor(rays_a, "rays_a", torch.int64, 2, 3)
    R, N, dev = rays_a.shape[0], ws.shape[0], ws.device
    loss = torch.zeros(R, dtype=torch.float32, device=dev)
    wsi = torch.zeros(N, dtype=torch.float32, device=dev)
    wtsi = torch.zeros(N, dtype=torch.float32, device=dev)
    call("arn_distortion_fw", ptr(ws), ptr(deltas), ptr(ts), ptr(rays_a), R, N, ptr(loss), ptr(wsi), ptr(wtsi), stream())
    return [loss, wsi, wtsi]


def distortion_loss_bw(dL_dloss, ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts, rays_a):
    """binding.cpp:212-231 -> losses.cu:143-181.  Returns dL_dws (N)."""
    for t, n in ((dL_dloss, "dL_dloss"), (ws_inclusive_scan, "ws_inclusive_scan"), (wts_inclusive_scan, "wts_inclusive_scan"),
                 (ws, "ws"), (deltas, "deltas"), (ts, "ts")):
        check_tensor(t, n, torch.float32, 1)
    check_tensor(rays_a, "rays_a", torch.int64, 2, 3)
    R, N = rays_a.shape[0], ws.shape[0]
    out = torch.zeros(N, dtype=torch.float32, device=ws.device)
    call("arn_distortion_bw", ptr(dL_dloss), ptr(ws_inclusive_scan), ptr(wts_inclusive_scan), ptr(ws), ptr(deltas),
         ptr(ts), ptr(rays_a), R, N, ptr(out), stream())
    return out


def segment_sums(dL_dxyzs, dL_ddirs, ts, rays_a):
    """Per-row sums for RayMarcher.backward (replaces torch_scatter.segment_csr, custom_functions.py:104-112)."""
    check_tensor(dL_dxyzs, "dL_dxyzs", torch.float32, 2, 3); check_tensor(ts, "ts", torch.float32, 1)
    check_tensor(rays_a, "rays_a", torch.int64, 2, 3)
    if dL_ddirs is not None:
        check_tensor(dL_ddirs, "dL_ddirs", torch.float32, 2, 3)
    R = rays_a.shape[0]
    d_o = torch.empty(R, 3, dtype=torch.float32, device=ts.device)
    d_d = torch.empty(R, 3, dtype=torch.float32, device=ts.device)
    call("arn_march_train_bw", ptr(dL_dxyzs), ptr(dL_ddirs), ptr(ts), ptr(rays_a), R, ptr(d_o), ptr(d_d), stream())
    return d_o, d_d
