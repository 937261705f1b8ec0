"""Autograd wrappers of the reference (models/custom_functions.py:8-173), same names, inputs and outputs, over
libarnerf.so instead of `vren` / torch_scatter."""
import torch
from torch.amp import custom_bwd, custom_fwd

from . import vren


class RayAABBIntersector(torch.autograd.Function):
    """custom_functions.py:8-29.  Returns (hits_cnt (R), hits_t (R,max_hits,2), hits_voxel_idx (R,max_hits)), near to far, -1 = no hit."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, center, half_size, max_hits):
        return tuple(vren.ray_aabb_intersect(rays_o, rays_d, center, half_size, max_hits))


class RaySphereIntersector(torch.autograd.Function):
    """custom_functions.py:32-52."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, center, radii, max_hits):
        return tuple(vren.ray_sphere_intersect(rays_o, rays_d, center, radii, max_hits))


class RayMarcher(torch.autograd.Function):
    """custom_functions.py:55-112.  Outputs rays_a (R,3) = (ray_idx, start_idx, N_samples), xyzs, dirs (N,3), deltas, ts (N),
    total_samples.  `noise` (optional, (R) float32 in [0,1)) replaces the internal torch.rand_like draw for parity tests."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, grid_size,
                max_samples, noise=None):
        if noise is None:
            noise = torch.rand_like(rays_o[:, 0])  # same RNG call as custom_functions.py:83
        rays_a, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(
            rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise.contiguous(), grid_size,
            max_samples)
        total_samples = counter[0]
        ctx.save_for_backward(rays_a, ts)
        ctx.mark_non_differentiable(rays_a, total_samples)
        return rays_a, xyzs, dirs, deltas, ts, total_samples

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dL_drays_a, dL_dxyzs, dL_ddirs, dL_ddeltas, dL_dts, dL_dtotal_samples):
        rays_a, ts = ctx.saved_tensors
        if dL_dxyzs is None:
            dL_dxyzs = torch.zeros(ts.shape[0], 3, device=ts.device)
        dL_drays_o, dL_drays_d = vren.segment_sums(dL_dxyzs.contiguous().float(),
                                                   None if dL_ddirs is None else dL_ddirs.contiguous().float(), ts, rays_a)
        return dL_drays_o, dL_drays_d, None, None, None, None, None, None, None, None


class VolumeRenderer(torch.autograd.Function):
    """custom_functions.py:115-159.  Outputs total_samples (scalar), opacity (R), depth (R), rgb (R,3), ws (N)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, ts, rays_a, T_threshold):
        total_samples, opacity, depth, rgb, ws = vren.composite_train_fw(
            sigmas.contiguous(), rgbs.contiguous(), deltas.contiguous(), ts.contiguous(), rays_a, T_threshold)
        ctx.save_for_backward(sigmas, rgbs, deltas, ts, rays_a, opacity, depth, rgb, ws)
        ctx.T_threshold = T_threshold
        ctx.set_materialize_grads(False)
        return total_samples.sum(), opacity, depth, rgb, ws

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dL_dtotal_samples, dL_dopacity, dL_ddepth, dL_drgb, dL_dws):
        sigmas, rgbs, deltas, ts, rays_a, opacity, depth, rgb, ws = ctx.saved_tensors
        z = lambda ref: torch.zeros_like(ref)
        dL_dopacity = z(opacity) if dL_dopacity is None else dL_dopacity.contiguous()
        dL_ddepth = z(depth) if dL_ddepth is None else dL_ddepth.contiguous()
        dL_drgb = z(rgb) if dL_drgb is None else dL_drgb.contiguous()
        dL_dws = None if dL_dws is None else dL_dws.contiguous()
        dL_dsigmas, dL_drgbs = vren.composite_train_bw(dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas.contiguous(),
                                                       rgbs.contiguous(), ws, deltas.contiguous(), ts.contiguous(), rays_a,
                                                       opacity, depth, rgb, ctx.T_threshold)
        return dL_dsigmas, dL_drgbs, None, None, None, None


class TruncExp(torch.autograd.Function):
    """custom_functions.py:162-173."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dL_dout):
        x = ctx.saved_tensors[0]
        return dL_dout * torch.exp(x.clamp(-15, 15))
