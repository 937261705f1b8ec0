"""Minimal torch-only training step driver for the hot path (stands in for train.py:53-198's LightningModule, which
needs pytorch-lightning / apex).  One process per GPU; rays are sharded across ranks, parameters replicated, and
hash-table + MLP gradients are summed with one NCCL all-reduce per step.

Step = density-grid update every 16 steps (train.py:175-178) -> render(train) -> NeRFLoss -> backward ->
gradient all-reduce -> fused Adam (lr 1e-2, eps 1e-15, train.py:146) with fp16 working-copy refresh.
"""
import ctypes as C
import math
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import FIELD_SCRATCH_BYTES, FieldWs, TrainCfg, call, ptr, stream
from .losses import NeRFLoss
from .rendering import MAX_SAMPLES, NEAR_DISTANCE, render


class FusedAdam:
    """apex FusedAdam(lr, eps=1e-15) replacement (train.py:146) on arn_adam_step: one pass does un-scale, Adam, the
    fp16 refresh of the working copy and the zeroing of the gradient buffer.

    world > 1: the gradient exchange lives here.  Small parameters are all-reduced and updated on every rank.  Parameters
    of at least `shard_min_numel` elements (the hash table) are SHARDED (sharding.py): reduce-scatter of the gradient,
    Adam on this rank's contiguous slice only, all-gather of the fp16 working copy -- 6 instead of 8 bytes per parameter
    on the wire and 1/world of Adam's HBM traffic.  The fp32 master of a sharded parameter is current only inside the
    owner's slice until gather_master() (call it before saving a checkpoint)."""

    def __init__(self, params_and_caches, lr=1e-2, betas=(0.9, 0.999), eps=1e-15, world=1, rank=0, shard_min_numel=1 << 20, exchange="p2p",
                 group_bounds=None):
        """group_bounds: element boundaries [0, b1, .., numel] of the LEVEL GROUPS of the large parameter (the hash table behind
        its 3072 MLP weights), matching arn_train_set_level_groups: step(group_events=) then updates / exchanges a group as soon
        as its gradient is final, beside the hash-grid backward of the next group."""
        from .sharding import PeerExchange, padded_numel, shard_size
        self.items = []
        self.world, self.rank = world, rank
        self.group_bounds = group_bounds
        self.opt_stream = self.opt_done = None
        for p, cache in params_and_caches:
            if p.numel() == 0:
                continue
            n = p.numel()
            sharded = world > 1 and n >= shard_min_numel and cache is not None
            if sharded:
                S, P = shard_size(n, world), padded_numel(n, world)
                lo = rank * S
                cnt = max(0, min(lo + S, n) - lo)
                px = None
                if exchange == "p2p" and dist.get_backend() == "nccl":
                    # peer-memory exchange fused with Adam (csrc/arn_p2p.cu); every rank must succeed, else all fall back to NCCL
                    try:
                        px = PeerExchange(n, world, rank, p.device, bounds=group_bounds if group_bounds and group_bounds[-1] == n else None)
                        ok = torch.ones(1, device=p.device)
                    except RuntimeError as e:
                        print(f"[ar_nerf_b200] peer-memory exchange unavailable on rank {rank}: {e}")
                        ok = torch.zeros(1, device=p.device)
                    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                    if float(ok.item()) == 0:
                        px = None
                if px is not None:
                    gpad = px.grad
                    cache.adopt(p, px.p16)
                    S = px.owned  # m, v: the rank's owned runs, group after group
                else:
                    gpad = torch.zeros(P, dtype=p.dtype, device=p.device)
                    cache.reserve(p, P)
                p.grad = gpad[:n].view_as(p)       # the kernels accumulate here; the tail pads the last shard
                st = dict(S=S, P=P, lo=lo, cnt=cnt, gpad=gpad, px=px,
                          gshard=None if px is not None else torch.empty(S, dtype=p.dtype, device=p.device))
                m, v = torch.zeros(S, dtype=p.dtype, device=p.device), torch.zeros(S, dtype=p.dtype, device=p.device)
            else:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                st = None
                m, v = torch.zeros_like(p), torch.zeros_like(p)
            self.items.append((p, cache, m, v, st))
        self.lr, self.betas, self.eps, self.t = lr, betas, eps, 0
        self.small_stream = self.small_done = None
        self.small_pending = False
        # the small all-reduces ride on their own communicator so that they overlap the table's reduce-scatter instead of
        # queueing behind it (a 28 KB all-reduce is pure latency: ~28 us at 8 GPUs)
        self.small_group = dist.new_group() if world > 1 and any(st is not None for *_, st in self.items) else None

    def step(self, inv_grad_scale=1.0, group_events=None):
        """group_events: one torch.cuda.Event per level group, recorded by the backward where the group's gradient is final
        (NGPTrainer arms them with arn_train_set_level_groups); None = everything behind the current stream's work."""
        from .sharding import all_gather_shards, reduce_scatter_sum
        self.t += 1
        s_ = stream()
        hyper = (float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), self.t, float(inv_grad_scale))
        if self.world == 1 and len(self.items) == 2:
            (pa, ca, ma, va, _), (pb, cb, mb, vb, _) = sorted(self.items, key=lambda it: -it[0].numel())
            ok = pa.numel() >= 4096 and pa.numel() % 4 == 0 and pb.numel() <= (1 << 20) and ca is not None and cb is not None
            if ok and group_events is not None and self.group_bounds and self.group_bounds[-1] == pa.numel():
                # single GPU, pipelined: Adam of a level group runs on the optimizer stream as soon as the group's gradient is
                # final, beside the hash-grid backward of the next group (HBM-bound against L2-reduction-bound); the colour
                # net's gradient is final with the first event
                if self.opt_stream is None:
                    self.opt_stream, self.opt_done = torch.cuda.Stream(device=pa.device, priority=-1), torch.cuda.Event()
                os_, oh = self.opt_stream, self.opt_stream.cuda_stream
                fa, ga, p16a = pa.data.view(-1), pa.grad.view(-1), ca.get(pa)
                for g, ev in enumerate(group_events):
                    lo, hi = self.group_bounds[g], self.group_bounds[g + 1]
                    os_.wait_event(ev)
                    if g == 0:
                        call("arn_adam_step2", ptr(fa[lo:hi]), ptr(ga[lo:hi]), ptr(ma[lo:hi]), ptr(va[lo:hi]), ptr(p16a[lo:hi]), hi - lo,
                             ptr(pb.data), ptr(pb.grad), ptr(mb), ptr(vb), ptr(cb.get(pb)), pb.numel(), *hyper, 1, oh)
                    else:
                        call("arn_adam_step", ptr(fa[lo:hi]), ptr(ga[lo:hi]), ptr(ma[lo:hi]), ptr(va[lo:hi]), ptr(p16a[lo:hi]), hi - lo, *hyper, 1, oh)
                self.opt_done.record(os_)
                torch.cuda.current_stream().wait_event(self.opt_done)
                ca.mark_fresh(pa); cb.mark_fresh(pb)
                return
            if ok:
                # single GPU: the hash table and the colour net in one launch (arn_adam_step2)
                call("arn_adam_step2", ptr(pa.data), ptr(pa.grad), ptr(ma), ptr(va), ptr(ca.get(pa)), pa.numel(),
                     ptr(pb.data), ptr(pb.grad), ptr(mb), ptr(vb), ptr(cb.get(pb)), pb.numel(), *hyper, 1, s_)
                ca.mark_fresh(pa); cb.mark_fresh(pb)
                return
        pending = {}
        if self.world > 1:
            for i, (p, cache, m, v, st) in enumerate(self.items):
                if st is None:
                    if group_events is not None:
                        # the MLP gradients are final behind the first group's event: their all-reduce starts there (a large
                        # parameter that is NOT sharded -- shard_optimizer=False -- is final behind the last one)
                        if self.small_stream is None:
                            self.small_stream, self.small_done = torch.cuda.Stream(device=p.device), torch.cuda.Event()
                        grouped = self.group_bounds and p.numel() == self.group_bounds[-1]
                        self.small_stream.wait_event(group_events[-1] if grouped else group_events[0])
                        with torch.cuda.stream(self.small_stream):
                            pending[i] = dist.all_reduce(p.grad, group=self.small_group, async_op=True)
                    else:
                        pending[i] = dist.all_reduce(p.grad, group=self.small_group, async_op=True)
        # sharded parameters first: their collectives are the long ones
        order = sorted(range(len(self.items)), key=lambda i: self.items[i][4] is None)
        for i in order:
            p, cache, m, v, st = self.items[i]
            p16 = cache.get(p) if cache is not None else None
            if st is None:
                if i in pending:
                    # small parameters (the colour net): their all-reduce and their Adam run on a side stream, beside the table's
                    # exchange instead of behind it; step() ends with the calling stream waiting for `small_done`
                    if self.small_stream is None:
                        self.small_stream, self.small_done = torch.cuda.Stream(device=p.device), torch.cuda.Event()
                    with torch.cuda.stream(self.small_stream):
                        pending[i].wait()
                        call("arn_adam_step", ptr(p.data), ptr(p.grad), ptr(m), ptr(v), ptr(p16), p.numel(), *hyper, 1, self.small_stream.cuda_stream)
                    self.small_done.record(self.small_stream)
                    self.small_pending = True
                else:
                    call("arn_adam_step", ptr(p.data), ptr(p.grad), ptr(m), ptr(v), ptr(p16), p.numel(), *hyper, 1, s_)
            elif st["px"] is not None:
                ev = group_events if (group_events is not None and len(group_events) == len(st["px"].groups)) else None
                st["px"].step(p.data.view(-1), m, v, hyper, self.t, s_, group_events=ev)
            else:
                reduce_scatter_sum(st["gpad"], st["gshard"], self.rank, self.world)
                st["gpad"].zero_()
                lo, cnt = st["lo"], st["cnt"]
                if cnt > 0:
                    flat = p.data.view(-1)
                    call("arn_adam_step", ptr(flat[lo:lo + cnt]), ptr(st["gshard"]), ptr(m), ptr(v), ptr(p16[lo:lo + cnt]), cnt, *hyper, 0, s_)
                all_gather_shards(p16[:st["P"]], self.rank, self.world)
            if cache is not None:
                cache.mark_fresh(p)
        self.wait_small()  # the stream that called step() sees the side-stream update: it ran beside the exchange, not behind it

    def wait_small(self):
        """Make the current stream wait for the small parameters' side-stream update (no-op if none is pending)."""
        if self.small_pending:
            torch.cuda.current_stream().wait_event(self.small_done)
            self.small_pending = False

    def pending_zero(self):
        """The peer exchange whose gradient buffer is still being zeroed on its side stream, or None."""
        for *_, st in self.items:
            if st is not None and st["px"] is not None and st["px"].zero_pending:
                return st["px"]
        return None

    def wait_zeroed(self):
        for *_, st in self.items:
            if st is not None and st["px"] is not None:
                st["px"].wait_zeroed()

    @torch.no_grad()
    def gather_master(self):
        """Make the fp32 master of every sharded parameter current on all ranks (checkpoints, evaluation in fp32)."""
        from .sharding import all_gather_shards
        for p, cache, m, v, st in self.items:
            if st is None:
                continue
            if st["px"] is not None:
                st["px"].gather_owned(p.data.view(-1))
                cache.mark_fresh(p)
                continue
            tmp = torch.zeros(st["P"], dtype=p.dtype, device=p.device)
            lo, cnt = st["lo"], st["cnt"]
            flat = p.data.view(-1)
            tmp[lo:lo + cnt] = flat[lo:lo + cnt]
            all_gather_shards(tmp, self.rank, self.world)
            flat.copy_(tmp[:p.numel()])
            cache.mark_fresh(p)


class _MarchSet:
    """What arn_train_march produces for one batch (and arn_train_fwbw_marched consumes).  Two sets alternate so that
    the next batch can be marched while the current one is still in its field / compositing kernels."""

    def __init__(self, n_rays, capacity, device):
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)
        self.rays_a = torch.empty(n_rays, 3, dtype=torch.int64, device=device)
        self.counter = torch.zeros(2, dtype=torch.int32, device=device)
        self.xyzs, self.dirs, self.deltas, self.ts = f(capacity, 3), f(capacity, 3), f(capacity), f(capacity)
        self.loss = torch.zeros(1, dtype=torch.float32, device=device)  # zeroed by the set's march, accumulated by its step
        self.inputs = None   # (rays_o, rays_d, noise) marched into this set: kept alive until the set is consumed
        self.ready = torch.cuda.Event()  # recorded after a side-stream (prefetched) march
        self.pending = False             # a prefetched march is in flight / waiting to be consumed
        self.grid_epoch = -1             # occupancy refresh count the set was marched against
        self.cfg = self.cfg_key = self.host = None


class _FusedWorkspace:
    """Device buffers of the fused step (arn_train_march + arn_train_fwbw_marched), allocated once: per-ray tensors for
    `n_rays`, per-sample tensors for `capacity` samples (n_rays * max_samples can never overflow)."""

    def __init__(self, n_rays, capacity, device):
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)
        h = lambda *s: torch.empty(*s, dtype=torch.float16, device=device)
        R, N = n_rays, capacity
        self.n_rays, self.capacity = R, N
        self.march = [_MarchSet(R, N, device), _MarchSet(R, N, device)]
        self.cur = 0
        # march scratch: shared by the two sets (marches are serialised: see NGPTrainer._march)
        self.hits_t = f(R, 1, 2)
        self.t_scratch = f(R * MAX_SAMPLES)
        self.count_scratch = torch.empty(R, dtype=torch.int32, device=device)
        self.total_samples = torch.empty(R, dtype=torch.int64, device=device)
        self.opacity, self.depth, self.rgb, self.rgb_final = f(R), f(R), f(R, 3), f(R, 3)
        self.dL_dopacity, self.dL_ddepth, self.dL_drgb = f(R), f(R), f(R, 3)
        self.sigmas, self.rgbs, self.ws_out = f(N), f(N, 3), f(N)
        self.dL_dsigmas, self.dL_drgbs, self.dfeat = f(N), f(N, 3), f(N, 32)
        self.feat, self.hid = h(N, 32), h(N, 64)  # no fp32 h: the tensor-core backward reads fp16(h) out of in32
        self.in32, self.hid1, self.hid2 = h(N, 32), h(N, 64), h(N, 64)
        self.wimg = torch.empty(FIELD_SCRATCH_BYTES, dtype=torch.uint8, device=device)
        self.marched = torch.cuda.Event()
        # the prefetched march runs here, filling issue slots the main stream leaves idle
        self.side = torch.cuda.Stream(device=device, priority=int(os.environ.get("ARN_SIDE_PRIO", "0")))
        self.fork_armed = False


class NGPTrainer:
    def __init__(self, model, lr=1e-2, num_epochs=30, steps_per_epoch=1000, loss_func='raw', depth_loss_w=0.0,
                 distortion_loss_w=0.0, exp_step_factor=None, random_bg=False, grad_scale=1.0,
                 update_interval=16, warmup_steps=256, fused=True, sample_capacity=None, shard_optimizer=True, exchange="p2p"):
        self.model = model
        # the fused native step covers the default training configuration (train.py defaults: 'raw' loss, no distortion
        # loss, fixed background); anything else runs the eager render() + autograd path
        self.fused = fused and loss_func == 'raw' and distortion_loss_w == 0 and not random_bg and model.rgb_act == 'Sigmoid'
        self.loss_func, self.depth_loss_w = loss_func, depth_loss_w
        self.sample_capacity = sample_capacity
        self._ws = None
        self.exp_step_factor = (1 / 256 if model.scale > 0.5 else 0.0) if exp_step_factor is None else exp_step_factor
        self.random_bg = random_bg
        self.loss = NeRFLoss(num_epochs, loss_func, model.scale, depth_loss_w, lambda_distortion=distortion_loss_w)
        self.grad_scale = grad_scale  # static outer loss scale (the reference uses PL's dynamic GradScaler)
        self.update_interval, self.warmup_steps = update_interval, warmup_steps
        self.base_lr, self.num_epochs, self.steps_per_epoch = lr, num_epochs, steps_per_epoch
        st = model.field_state
        st.direct_grad = True
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        # Level-major hash-grid backward (arn_train_set_level_groups): the optimizer / the multi-GPU exchange of a level group can
        # start where that group's gradient is final, beside the backward of the next group.  OFF by default ("0,16"): measured
        # on B200 the split costs more than the overlap returns -- the backward's levels overlap inside ONE launch (16 single-
        # level launches take 438 us, the fused launch 79 us), so a second launch adds ~29 us: 8 GPUs 0.429 ms per step
        # unsplit, 0.462 with "0,11,16"; 2 GPUs 0.376 / 0.405; 1 GPU (Adam per group) 0.332 / 0.352.
        # ARN_LEVEL_GROUPS="0,11,16" switches it on.
        default_groups = "0,16"
        lv = [int(x) for x in os.environ.get("ARN_LEVEL_GROUPS", default_groups).split(",")]
        self.level_groups = lv if (self.fused and len(lv) > 2 and lv[0] == 0 and lv[-1] == 16) else None
        bounds = None
        if self.level_groups:
            off = model.geometry.offset
            n_xyz = model.xyz_encoder.params.numel()
            bounds = [0] + [3072 + 2 * int(off[l]) for l in self.level_groups[1:-1]] + [n_xyz]
            self._lg_begin = (C.c_int * len(lv))(*lv)
            self._lg_events = [torch.cuda.Event() for _ in range(len(lv) - 1)]
            for e in self._lg_events:
                e.record()  # creates the CUDA event handles
            self._lg_handles = (C.c_void_p * (len(lv) - 1))(*[e.cuda_event for e in self._lg_events])
        # every parameter of the model is optimised, as train.py:141-146 does (all named parameters except the pose refinement's
        # dR / dT): in HDR mode that includes the three tonemapper nets (no fp16 working copy: they are evaluated by torch)
        main_params = {id(model.xyz_encoder.params), id(model.rgb_net.params)}
        extra = [(p, None) for n_, p in model.named_parameters() if id(p) not in main_params and n_ not in ('dR', 'dT')]
        self.opt = FusedAdam([(model.xyz_encoder.params, st.cache_xyz), (model.rgb_net.params, st.cache_rgb)] + extra, lr,
                             world=self.world, rank=self.rank, shard_min_numel=(1 << 20) if shard_optimizer else (1 << 62),
                             exchange=exchange, group_bounds=bounds)
        self.global_step = 0
        self._grid_epoch = 0
        self.fork_stage = int(os.environ.get("ARN_FORK_STAGE", "2"))  # where the prefetched march joins the step (arn_train_set_fork):
        # measured on B200, ms per step with the fork at stage 0..4: 0.375 / 0.358 / 0.345 / 0.356 / 0.390 -- behind compositing the
        # march shares the SMs with the MLP backward (8 warps per SM, latency-bound) and the memory-bound hash-grid backward + Adam

    def lr_at(self, step):
        """CosineAnnealingLR(T_max=num_epochs, eta_min=lr/30) stepped once per epoch (train.py:150-152)."""
        e = step // self.steps_per_epoch
        eta_min = self.base_lr / 30
        return eta_min + (self.base_lr - eta_min) * (1 + math.cos(math.pi * e / self.num_epochs)) / 2

    def _workspace(self, R, dev):
        if self._ws is None or self._ws.n_rays != R:
            self._ws = _FusedWorkspace(R, (int(self.sample_capacity or R * MAX_SAMPLES) + 127) // 128 * 128, dev)
        return self._ws

    def _cfg(self, ms, rgb_target):
        """arn_train_t for march set `ms` (its inputs must be set).  The struct is built once per set and only the batch
        pointers are refreshed per call (the workspace, the fp16 parameter copies and the gradients never move)."""
        m, st, w = self.model, self.model.field_state, self._ws
        ro, rd, nz = ms.inputs
        p16x = st.cache_xyz.get(m.xyz_encoder.params); p16c = st.cache_rgb.get(m.rgb_net.params)
        key = (id(w), m.density_bitfield.data_ptr(), p16x.data_ptr(), p16c.data_ptr(), m.xyz_encoder.params.grad.data_ptr(),
               m.rgb_net.params.grad.data_ptr(), self.exp_step_factor, self.grad_scale, st.loss_scale)
        if ms.cfg is None or ms.cfg_key != key:
            center, half = m.host_box()
            bgv = 1.0 if self.exp_step_factor == 0 else 0.0
            ms.host = ((C.c_float * 3)(*center), (C.c_float * 3)(*half), (C.c_float * 3)(bgv, bgv, bgv))  # kept alive with the struct
            cast = lambda a: C.cast(a, C.c_void_p)
            ms.cfg = TrainCfg(
                None, None, None, None, ro.shape[0],
                ptr(m.density_bitfield), m.cascades, m.grid_size, float(m.scale), float(self.exp_step_factor), MAX_SAMPLES, 1e-4, NEAR_DISTANCE,
                cast(ms.host[0]), cast(ms.host[1]), cast(st.mn), cast(st.mx),
                st.geometry.c_levels, ptr(p16x), ptr(p16c), st.rgb_act,
                cast(ms.host[2]), float(self.loss.lambda_opacity), float(self.depth_loss_w), float(self.grad_scale), float(st.loss_scale),
                ptr(w.hits_t), ptr(ms.rays_a), ptr(ms.counter), ptr(w.t_scratch), ptr(w.count_scratch), ptr(w.total_samples),
                ptr(w.opacity), ptr(w.depth), ptr(w.rgb), ptr(w.rgb_final), ptr(w.dL_dopacity), ptr(w.dL_ddepth), ptr(w.dL_drgb),
                w.capacity, ptr(ms.xyzs), ptr(ms.dirs), ptr(ms.deltas), ptr(ms.ts), ptr(w.sigmas), ptr(w.rgbs), ptr(w.ws_out),
                ptr(w.dL_dsigmas), ptr(w.dL_drgbs), ptr(w.dfeat),
                FieldWs(ptr(w.feat), ptr(w.hid), None, ptr(w.in32), ptr(w.hid1), ptr(w.hid2), ptr(w.wimg)),
                ptr(m.xyz_encoder.params.grad), ptr(m.rgb_net.params.grad), ptr(ms.loss))
            ms.cfg_key = key
        c = ms.cfg
        c.rays_o, c.rays_d, c.noise = ro.data_ptr(), rd.data_ptr(), nz.data_ptr()
        c.rgb_target = None if rgb_target is None else rgb_target.data_ptr()
        return c

    def _march(self, ms, rays_o, rays_d, noise, cuda_stream):
        """Geometry half of the step for one batch into march set `ms` on `cuda_stream` (raw handle)."""
        R, dev = rays_o.shape[0], rays_o.device
        if noise is None:
            noise = torch.rand(R, device=dev)  # the draw RayMarcher.forward makes (custom_functions.py:83)
        ms.inputs = (rays_o.contiguous().float(), rays_d.contiguous().float(), noise.contiguous().float())
        call("arn_train_march", C.byref(self._cfg(ms, None)), cuda_stream)

    def _fused_fwbw(self, rays_o, rays_d, rgb_target, noise, next_rays=None, next_is_update=False):
        R, dev = rays_o.shape[0], rays_o.device
        w = self._workspace(R, dev)
        main = torch.cuda.current_stream()
        main_h = main.cuda_stream
        ms = w.march[w.cur]
        pre = ms.pending and ms.inputs[0] is rays_o and ms.inputs[1] is rays_d and ms.grid_epoch == self._grid_epoch
        if ms.pending:
            main.wait_event(ms.ready)      # marched ahead on the side stream while the previous step was running
            ms.pending = False
        if not pre:
            self._march(ms, rays_o, rays_d, noise, main_h)
        prefetch = None
        if next_rays is not None and not next_is_update:
            n_ro, n_rd = next_rays[0], next_rays[1]
            if n_ro.is_contiguous() and n_rd.is_contiguous() and n_ro.dtype == torch.float32 and n_rd.dtype == torch.float32 and n_ro.shape[0] == R:
                prefetch = (n_ro, n_rd, next_rays[2] if len(next_rays) > 2 else None)
        self._keep_target = rgb_target.contiguous().float()
        if prefetch is not None:
            # the side stream forks where the library records `marched`: from there on the shared march scratch is free and
            # the next batch's rays exist (any stage of this call qualifies); see arn_train_set_fork for the choice of stage
            if not w.fork_armed:
                w.marched.record(main)     # creates the CUDA event handle
                w.fork_armed = True
            call("arn_train_set_fork", self.fork_stage, C.c_void_p(w.marched.cuda_event))
        px = self.opt.pending_zero()
        if px is not None:  # the gradient buffer is being zeroed on a side stream: the MLP backward (first writer) waits for it
            call("arn_train_set_join", 2, C.c_void_p(px.zeroed.cuda_event))
        if self.level_groups:
            call("arn_train_set_level_groups", len(self._lg_events), self._lg_begin, self._lg_handles)
        try:
            call("arn_train_fwbw_marched", C.byref(self._cfg(ms, self._keep_target)), main_h)
        finally:
            if self.level_groups:
                call("arn_train_set_level_groups", 0, None, None)
        if px is not None:
            call("arn_train_set_join", 0, None)
            px.zero_pending = False
        if prefetch is not None:
            call("arn_train_set_fork", 0, None)
        if prefetch is not None:
            # geometry of the next batch, concurrently with this batch's field / compositing / optimizer kernels: it reads the
            # rays and the occupancy bits only.  (Skipped when the next step refreshes the occupancy grid first.)
            nxt = w.march[w.cur ^ 1]
            w.side.wait_event(w.marched)
            n_ro, n_rd, n_noise = prefetch
            if n_noise is None:
                # next in the RNG stream, exactly where the next step would draw it; generated on the side stream (the Philox
                # offset advances on the host, the values do not depend on the stream) so that nothing queues behind this step
                with torch.cuda.stream(w.side):
                    n_noise = torch.rand(R, device=dev)
            self._march(nxt, n_ro, n_rd, n_noise, w.side.cuda_stream)
            nxt.ready.record(w.side)
            nxt.pending = True
            nxt.grid_epoch = self._grid_epoch
        w.cur ^= 1
        # everything returned is a VIEW into the reused workspace: the per-ray outputs are valid until the next step, the loss
        # and the march products (rm_samples, rays_a, ts/deltas) until the step after it; per-sample buffers hold
        # rm_samples samples.  Clone what has to live longer.
        results = {'rgb': w.rgb_final, 'opacity': w.opacity, 'depth': w.depth, 'rm_samples': ms.counter[0],
                   'rays_a': ms.rays_a, 'total_samples_per_ray': w.total_samples, 'ts_buf': ms.ts, 'deltas_buf': ms.deltas,
                   'ws_buf': w.ws_out}
        return ms.loss[0], results

    def train_step(self, rays_o, rays_d, rgb_target, noise=None, update_grid=True, next_rays=None):
        """One optimisation step.  next_rays = (rays_o, rays_d[, noise]) of the FOLLOWING call, if the caller already has
        them on the device: their ray march is then overlapped with this step (pass the very same tensors next time)."""
        m = self.model
        if update_grid and self.global_step % self.update_interval == 0:
            # (the refresh after this one is a steady-state refresh as soon as it lies behind the warm-up: its cell selection
            # is then computed ahead, under the steps in between -- NGP.update_density_grid)
            m.update_density_grid(0.01 * MAX_SAMPLES / 3 ** 0.5, warmup=self.global_step < self.warmup_steps,
                                  prefetch_next=self.global_step + self.update_interval >= self.warmup_steps)
            self._grid_epoch += 1
        if self.fused:
            next_is_update = update_grid and (self.global_step + 1) % self.update_interval == 0
            loss, results = self._fused_fwbw(rays_o, rays_d, rgb_target, noise, next_rays, next_is_update)
            self.opt.lr = self.lr_at(self.global_step)
            pipelined = self.level_groups and not os.environ.get("ARN_NO_PIPELINED_OPT")  # A/B: split backward, optimizer behind it as before
            self.opt.step(inv_grad_scale=1.0 / (self.grad_scale * self.world), group_events=self._lg_events if pipelined else None)
            self.global_step += 1
            return loss, results
        self.opt.wait_zeroed()
        kwargs = {'test_time': False, 'random_bg': self.random_bg, 'exp_step_factor': self.exp_step_factor}
        if noise is not None:
            kwargs['noise'] = noise
        results = render(m, rays_o, rays_d, **kwargs)
        loss_d = self.loss(results, {'rgb': rgb_target})
        loss = sum(lo.mean() for lo in loss_d.values())
        (loss * self.grad_scale).backward()
        self.opt.lr = self.lr_at(self.global_step)
        self.opt.step(inv_grad_scale=1.0 / (self.grad_scale * self.world))
        self.global_step += 1
        return loss.detach(), results


class DeviceDataset:
    """The training set resident in HBM: poses (n_images,3,4), images (n_images, H*W, C) f32 -- the reference's `dataset.rays`,
    which datasets/base.py:32 indexes on the HOST -- and the pixel directions (H*W,3) (`self.directions`, train.py:78)."""

    def __init__(self, poses, images, directions, device):
        self.poses = poses.to(device).float().contiguous()
        self.images = images.to(device).float().contiguous()
        self.directions = directions.to(device).float().contiguous()
        self.device = device


class BatchFeeder:
    """Host-drawn index batches -> device batches, two steps ahead of the training step.

    The reference's dataset draws `img_idxs` / `pix_idxs` on the host (datasets/base.py:24-30), gathers the pixels on the host
    and ships them; poses / directions are gathered on the device (train.py:121-126).  Here one step's host -> device traffic
    is the two index vectors (16 bytes per ray, ONE pinned block, ONE cudaMemcpyAsync on a copy stream) and arn_gather_batch
    produces (rays_o, rays_d, rgb) on that same stream into one of `depth` slots; events order slot reuse against the step
    that consumed the slot.  The trainer recognises a batch marched ahead by tensor identity, so get() returns the slot's
    own tensors."""

    def __init__(self, dataset, batch_size, depth=3):
        dev = dataset.device
        self.ds, self.B, self.depth = dataset, batch_size, depth
        ch = dataset.images.shape[2]
        self.idx = [torch.empty(2, batch_size, dtype=torch.int64, device=dev) for _ in range(depth)]
        self.slots = [(torch.empty(batch_size, 3, device=dev), torch.empty(batch_size, 3, device=dev), torch.empty(batch_size, ch, device=dev))
                      for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.landed, self.consumed = {}, {}
        self.h2d_bytes_per_batch = 2 * batch_size * 8

    def stage(self, i, idx_pinned):
        """Queue batch i: idx_pinned (2, B) int64 pinned = (img_idxs, pix_idxs).  No-op if already staged."""
        from . import vren
        if i in self.landed:
            return
        k = i % self.depth
        with torch.cuda.stream(self.copy_stream):
            if i - self.depth in self.consumed:
                self.copy_stream.wait_event(self.consumed[i - self.depth])  # the step that read this slot last has been queued and ordered
            self.idx[k].copy_(idx_pinned, non_blocking=True)
            vren.gather_batch(self.ds.poses, self.idx[k][0], self.idx[k][1], self.ds.images, directions=self.ds.directions, out=self.slots[k])
            ev = torch.cuda.Event(); ev.record(self.copy_stream)
            self.landed[i] = ev

    def get(self, i):
        """(rays_o, rays_d, rgb) of batch i; the current stream waits for its copy + gather."""
        torch.cuda.current_stream().wait_event(self.landed[i])
        return self.slots[i % self.depth]

    def done(self, i):
        """Call after the step that consumes batch i has been queued on the current stream."""
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream())
        self.consumed[i] = ev
        self.consumed.pop(i - 2 * self.depth, None); self.landed.pop(i - self.depth, None)
