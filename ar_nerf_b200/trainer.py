"""Minimal torch-only training step driver for the hot path (stands in for train.py:53-198's LightningModule, which
needs pytorch-lightning / apex).  One process per GPU; rays are sharded across ranks, parameters replicated, and
hash-table + MLP gradients are summed with one NCCL all-reduce per step.

Step = density-grid update every 16 steps (train.py:175-178) -> render(train) -> NeRFLoss -> backward ->
gradient all-reduce -> fused Adam (lr 1e-2, eps 1e-15, train.py:146) with fp16 working-copy refresh.
"""
import math

import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, ptr, stream
from .losses import NeRFLoss
from .rendering import MAX_SAMPLES, render


class FusedAdam:
    """apex FusedAdam(lr, eps=1e-15) replacement (train.py:146) on arn_adam_step: one pass does un-scale, Adam, the
    fp16 refresh of the working copy and the zeroing of the gradient buffer."""

    def __init__(self, params_and_caches, lr=1e-2, betas=(0.9, 0.999), eps=1e-15):
        self.items = []
        for p, cache in params_and_caches:
            if p.numel() == 0:
                continue
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            self.items.append((p, cache, torch.zeros_like(p), torch.zeros_like(p)))
        self.lr, self.betas, self.eps, self.t = lr, betas, eps, 0

    def step(self, inv_grad_scale=1.0):
        self.t += 1
        for p, cache, m, v in self.items:
            p16 = cache.get(p) if cache is not None else None
            call("arn_adam_step", ptr(p.data), ptr(p.grad), ptr(m), ptr(v), ptr(p16), p.numel(), float(self.lr),
                 float(self.betas[0]), float(self.betas[1]), float(self.eps), self.t, float(inv_grad_scale), 1, stream())
            if cache is not None:
                cache.mark_fresh(p)


class NGPTrainer:
    def __init__(self, model, lr=1e-2, num_epochs=30, steps_per_epoch=1000, loss_func='raw', depth_loss_w=0.0,
                 distortion_loss_w=0.0, exp_step_factor=None, random_bg=False, grad_scale=128.0,
                 update_interval=16, warmup_steps=256):
        self.model = model
        self.exp_step_factor = (1 / 256 if model.scale > 0.5 else 0.0) if exp_step_factor is None else exp_step_factor
        self.random_bg = random_bg
        self.loss = NeRFLoss(num_epochs, loss_func, model.scale, depth_loss_w, lambda_distortion=distortion_loss_w)
        self.grad_scale = grad_scale  # static outer loss scale (the reference uses PL's dynamic GradScaler)
        self.update_interval, self.warmup_steps = update_interval, warmup_steps
        self.base_lr, self.num_epochs, self.steps_per_epoch = lr, num_epochs, steps_per_epoch
        st = model.field_state
        st.direct_grad = True
        self.opt = FusedAdam([(model.xyz_encoder.params, st.cache_xyz), (model.rgb_net.params, st.cache_rgb)], lr)
        self.global_step = 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    def lr_at(self, step):
        """CosineAnnealingLR(T_max=num_epochs, eta_min=lr/30) stepped once per epoch (train.py:150-152)."""
        e = step // self.steps_per_epoch
        eta_min = self.base_lr / 30
        return eta_min + (self.base_lr - eta_min) * (1 + math.cos(math.pi * e / self.num_epochs)) / 2

    def train_step(self, rays_o, rays_d, rgb_target, noise=None, update_grid=True):
        m = self.model
        if update_grid and self.global_step % self.update_interval == 0:
            m.update_density_grid(0.01 * MAX_SAMPLES / 3 ** 0.5, warmup=self.global_step < self.warmup_steps)
        kwargs = {'test_time': False, 'random_bg': self.random_bg, 'exp_step_factor': self.exp_step_factor}
        if noise is not None:
            kwargs['noise'] = noise
        results = render(m, rays_o, rays_d, **kwargs)
        loss_d = self.loss(results, {'rgb': rgb_target})
        loss = sum(lo.mean() for lo in loss_d.values())
        (loss * self.grad_scale).backward()
        if self.world > 1:
            for p, _, _, _ in self.opt.items:
                dist.all_reduce(p.grad)
        self.opt.lr = self.lr_at(self.global_step)
        self.opt.step(inv_grad_scale=1.0 / (self.grad_scale * self.world))
        self.global_step += 1
        return loss.detach(), results
