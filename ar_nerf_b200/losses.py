"""losses.py of the reference (DistortionLoss :7-38, NeRFLoss :41-82), over libarnerf.so."""
import torch
from torch import nn

from . import vren


class DistortionLoss(torch.autograd.Function):
    """Mip-NeRF 360 distortion loss, DVGO-v2 formulation (losses.py:7-38).  Inputs ws, deltas, ts (N), rays_a (R,3) -> loss (R)."""

    @staticmethod
    def forward(ctx, ws, deltas, ts, rays_a):
        loss, ws_inclusive_scan, wts_inclusive_scan = vren.distortion_loss_fw(ws.contiguous(), deltas, ts, rays_a)
        ctx.save_for_backward(ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts, rays_a)
        return loss

    @staticmethod
    def backward(ctx, dL_dloss):
        ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts, rays_a = ctx.saved_tensors
        dL_dws = vren.distortion_loss_bw(dL_dloss.contiguous(), ws_inclusive_scan, wts_inclusive_scan, ws.contiguous(),
                                         deltas, ts, rays_a)
        return dL_dws, None, None, None


class NeRFLoss(nn.Module):
    """losses.py:41-82 (same constructor and result keys)."""

    def __init__(self, epoch, loss_set, grid_scale, lambda_depth, lambda_opacity=1e-3, lambda_distortion=1e-3):
        super().__init__()
        self.num_epoch = epoch
        self.grid_scale = grid_scale
        self.lambda_opacity = lambda_opacity
        self.lambda_depth = lambda_depth
        self.lambda_distortion = lambda_distortion
        if loss_set == 'raw':
            self.rgb_loss = lambda x_est, x_gt: (x_est - x_gt) / (x_est.detach() + 1e-3)
        elif loss_set == 'log':
            self.rgb_loss = lambda x_est, x_gt: torch.log((0.2935 + x_est) / (0.2935 + x_gt)) * 0.7607
        elif loss_set == 'tanh':
            self.rgb_loss = lambda x_est, x_gt: torch.tanh(x_est) - torch.tanh(x_gt)
        else:
            print('Unknown loss function!')

    def forward(self, results, target, **kwargs):
        d = {}
        d['rgb'] = self.rgb_loss(results['rgb'], target['rgb']) ** 2
        o = results['opacity'] + 1e-10
        d['opacity'] = self.lambda_opacity * (-o * torch.log(o))
        d['depth'] = -self.lambda_depth * torch.log((results['depth'] / self.grid_scale + 1e-10).clip(max=1.0))
        if self.lambda_distortion > 0:
            d['distortion'] = self.lambda_distortion * DistortionLoss.apply(results['ws'], results['deltas'],
                                                                            results['ts'], results['rays_a'])
        return d
