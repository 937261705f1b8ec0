"""The reference's autograd entry points (models/custom_functions.py:8-173) -- same class names, argument order and outputs --
over libarnerf.so instead of `vren` / torch_scatter."""
import torch
from torch.amp import custom_bwd, custom_fwd

from . import vren

_fp32_in = custom_fwd(device_type="cuda", cast_inputs=torch.float32)  # the reference's @custom_fwd(cast_inputs=torch.float32)
_bwd = custom_bwd(device_type="cuda")


def _dense(grad, like):
    """A missing (None) or strided incoming gradient as a contiguous tensor shaped like `like`."""
    return torch.zeros_like(like) if grad is None else grad.contiguous()


class RayAABBIntersector(torch.autograd.Function):
    """custom_functions.py:8-29: rays against boxes (center, half_size).  -> (hits_cnt (R), hits_t (R,max_hits,2),
    hits_voxel_idx (R,max_hits)), near to far, -1 where a ray has no (further) hit."""

    @staticmethod
    @_fp32_in
    def forward(ctx, origins, directions, center, half_size, max_hits):
        return tuple(vren.ray_aabb_intersect(origins, directions, center, half_size, max_hits))


class RaySphereIntersector(torch.autograd.Function):
    """custom_functions.py:32-52: the same against spheres (center, radii)."""

    @staticmethod
    @_fp32_in
    def forward(ctx, origins, directions, center, radii, max_hits):
        return tuple(vren.ray_sphere_intersect(origins, directions, center, radii, max_hits))


class RayMarcher(torch.autograd.Function):
    """custom_functions.py:55-112.  -> rays_a (R,3) = (ray_idx, start_idx, N_samples), xyzs, dirs (N,3), deltas, ts (N),
    total_samples.  `noise` (optional, (R) float32 in [0,1)) replaces the internal random draw (parity tests)."""

    @staticmethod
    @_fp32_in
    def forward(ctx, origins, directions, hits_t, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples, noise=None):
        jitter = torch.rand_like(origins[:, 0]) if noise is None else noise   # the RNG call of custom_functions.py:83
        rays_a, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(
            origins, directions, hits_t, density_bitfield, cascades, scale, exp_step_factor, jitter.contiguous(), grid_size, max_samples)
        n_marched = counter[0]
        ctx.save_for_backward(rays_a, ts)
        ctx.mark_non_differentiable(rays_a, n_marched)
        return rays_a, xyzs, dirs, deltas, ts, n_marched

    @staticmethod
    @_bwd
    def backward(ctx, _g_rays_a, g_xyzs, g_dirs, _g_deltas, _g_ts, _g_total):
        # xyz = o + t d, dir = d: a ray's gradients are sums over its samples (the reference's two segment_csr calls, :108-110)
        rays_a, ts = ctx.saved_tensors
        if g_xyzs is None:
            g_xyzs = ts.new_zeros(ts.shape[0], 3)
        g_o, g_d = vren.segment_sums(g_xyzs.contiguous().float(), None if g_dirs is None else g_dirs.contiguous().float(), ts, rays_a)
        return (g_o, g_d) + (None,) * 8


class VolumeRenderer(torch.autograd.Function):
    """custom_functions.py:115-159.  -> total_samples (scalar), opacity (R), depth (R), rgb (R,3), ws (N)."""

    @staticmethod
    @_fp32_in
    def forward(ctx, sigmas, rgbs, deltas, ts, rays_a, T_threshold):
        sigmas, rgbs, deltas, ts = (t.contiguous() for t in (sigmas, rgbs, deltas, ts))
        used, opacity, depth, rgb, ws = vren.composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T_threshold)
        ctx.save_for_backward(sigmas, rgbs, deltas, ts, rays_a, opacity, depth, rgb, ws)
        ctx.T_threshold = T_threshold
        ctx.set_materialize_grads(False)
        return used.sum(), opacity, depth, rgb, ws

    @staticmethod
    @_bwd
    def backward(ctx, _g_total, g_opacity, g_depth, g_rgb, g_ws):
        sigmas, rgbs, deltas, ts, rays_a, opacity, depth, rgb, ws = ctx.saved_tensors
        g_sigmas, g_rgbs = vren.composite_train_bw(
            _dense(g_opacity, opacity), _dense(g_depth, depth), _dense(g_rgb, rgb), None if g_ws is None else g_ws.contiguous(),
            sigmas, rgbs, ws, deltas, ts, rays_a, opacity, depth, rgb, ctx.T_threshold)
        return g_sigmas, g_rgbs, None, None, None, None


class TruncExp(torch.autograd.Function):
    """custom_functions.py:162-173: exp whose derivative is taken at the argument clamped to [-15, 15]."""

    @staticmethod
    @_fp32_in
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return x.exp()

    @staticmethod
    @_bwd
    def backward(ctx, g_out):
        (x,) = ctx.saved_tensors
        return g_out * x.clamp(-15, 15).exp()
