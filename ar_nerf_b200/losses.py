"""The reference's losses (losses.py: DistortionLoss :7-38, NeRFLoss :41-82) -- same class names, constructor arguments and
result keys -- over libarnerf.so."""
import torch
from torch import nn

from . import vren

# per-channel residual whose square is the colour loss (losses.py:49-52), by `loss_set`
_RESIDUAL = {
    'raw': lambda est, gt: (est - gt) / (est.detach() + 1e-3),
    'log': lambda est, gt: 0.7607 * torch.log((est + 0.2935) / (gt + 0.2935)),
    'tanh': lambda est, gt: est.tanh() - gt.tanh(),
}


class DistortionLoss(torch.autograd.Function):
    """Mip-NeRF 360 distortion loss in DVGO-v2's prefix-sum form (losses.py:7-38): per-sample weights, interval lengths and
    midpoints (N each) + rays_a (R,3) -> one loss value per ray (R)."""

    @staticmethod
    def forward(ctx, weights, intervals, midpoints, rays_a):
        weights = weights.contiguous()
        per_ray, w_prefix, wt_prefix = vren.distortion_loss_fw(weights, intervals, midpoints, rays_a)
        ctx.save_for_backward(w_prefix, wt_prefix, weights, intervals, midpoints, rays_a)
        return per_ray

    @staticmethod
    def backward(ctx, grad_per_ray):
        w_prefix, wt_prefix, weights, intervals, midpoints, rays_a = ctx.saved_tensors
        grad_weights = vren.distortion_loss_bw(grad_per_ray.contiguous(), w_prefix, wt_prefix, weights, intervals, midpoints, rays_a)
        return grad_weights, None, None, None


class NeRFLoss(nn.Module):
    """losses.py:41-82: forward(results, target) -> {'rgb', 'opacity', 'depth'[, 'distortion']}, unreduced."""

    def __init__(self, epoch, loss_set, grid_scale, lambda_depth, lambda_opacity=1e-3, lambda_distortion=1e-3):
        super().__init__()
        self.num_epoch, self.grid_scale = epoch, grid_scale
        self.lambda_depth, self.lambda_opacity, self.lambda_distortion = lambda_depth, lambda_opacity, lambda_distortion
        if loss_set in _RESIDUAL:
            self.rgb_loss = _RESIDUAL[loss_set]
        else:  # the reference only reports it (and fails at the first forward)
            print('Unknown loss function!')

    def forward(self, results, target, **kwargs):
        opacity = results['opacity'] + 1e-10
        terms = {
            'rgb': self.rgb_loss(results['rgb'], target['rgb']).square(),
            'opacity': self.lambda_opacity * -(opacity * opacity.log()),            # entropy: pushes opacities to 0 or 1
            'depth': -self.lambda_depth * (results['depth'] / self.grid_scale + 1e-10).clamp(max=1.0).log(),
        }
        if self.lambda_distortion > 0:
            terms['distortion'] = self.lambda_distortion * DistortionLoss.apply(results['ws'], results['deltas'], results['ts'], results['rays_a'])
        return terms
