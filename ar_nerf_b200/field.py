"""The field behind models/networks.py NGP: hash grid + density MLP + SH-4 + colour MLP on libarnerf.so.

Replaces the three tiny-cuda-nn modules the reference builds (networks.py:37-57 NetworkWithInputEncoding,
:59-66 Encoding, :68-78 Network).  The parameter containers keep tiny-cuda-nn's state-dict layout
(`xyz_encoder.params` = [3072 MLP weights | 2*entries table] fp32, `dir_encoder.params` empty,
`rgb_net.params` = 7168 fp32) so reference checkpoints load.
"""
import ctypes as C
import math

import numpy as np
import torch
from torch import nn
from torch.amp import custom_bwd, custom_fwd

from . import _lib
from ._lib import FIELD_SCRATCH_BYTES, FieldWs, Levels, call, ptr, stream

N_LEVELS = 16
DENSITY_MLP_PARAMS = 3072
RGB_MLP_PARAMS = 7168
DEFAULT_LOSS_SCALE = 128.0  # tiny-cuda-nn's default loss scale for fp16 gradients


class HashGeometry:
    """Level table in strict float32, computed ONCE on the host by libarnerf.so (arn_hashgrid_geometry) and read by
    every kernel (never recomputed on the device; SURVEY Appendix A.2 rounding hazard)."""

    def __init__(self, n_levels=16, base_resolution=16, per_level_scale=1.3195079, log2_hashmap_size=19):
        assert n_levels == N_LEVELS, "the kernels are specialised for L=16 (networks.py:33)"
        self.n_levels, self.base_resolution = n_levels, base_resolution
        self.per_level_scale, self.log2_hashmap_size = float(per_level_scale), log2_hashmap_size
        self.scale = np.zeros(n_levels, np.float32); self.res = np.zeros(n_levels, np.uint32)
        self.size = np.zeros(n_levels, np.uint32); self.offset = np.zeros(n_levels + 1, np.uint32)
        call("arn_hashgrid_geometry", n_levels, base_resolution, C.c_float(per_level_scale), log2_hashmap_size,
             self.scale.ctypes.data, self.res.ctypes.data, self.size.ctypes.data, self.offset.ctypes.data)
        self._finish()

    def _finish(self):
        self.total = int(self.offset[-1])
        self.c_levels = Levels(self.scale.ctypes.data, self.res.ctypes.data, self.size.ctypes.data, self.offset.ctypes.data)

    @classmethod
    def in_double(cls, n_levels=16, base_resolution=16, per_level_scale=1.3195079, log2_hashmap_size=19):
        """The same rule evaluated in float64: a tiny-cuda-nn build that computes the level scales in
        double lands just BELOW the integer where float32 lands just above it (scale 0.5, level 5: res 64 instead of 65), which
        shifts every later offset.  Used to read checkpoints written by such a build."""
        g = cls.__new__(cls)
        g.n_levels, g.base_resolution = n_levels, base_resolution
        g.per_level_scale, g.log2_hashmap_size = float(per_level_scale), log2_hashmap_size
        g.scale = np.zeros(n_levels, np.float32); g.res = np.zeros(n_levels, np.uint32)
        g.size = np.zeros(n_levels, np.uint32); g.offset = np.zeros(n_levels + 1, np.uint32)
        off = 0
        for l in range(n_levels):
            sc = 2.0 ** (l * math.log2(per_level_scale)) * base_resolution - 1.0
            res = int(math.ceil(sc)) + 1
            params = min(res ** 3, 0xffffffff // 2)
            params = (params + 7) // 8 * 8
            params = min(params, 1 << log2_hashmap_size)
            g.scale[l], g.res[l], g.size[l], g.offset[l] = np.float32(sc), res, params, off
            off += params
        g.offset[n_levels] = off
        g._finish()
        return g


def _xavier_uniform_(t, fan_out, fan_in, gen):
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    t.copy_((torch.rand(t.shape, generator=gen) * 2 - 1) * bound)


class NetworkWithInputEncoding(nn.Module):
    """Parameter container of the hash grid + density MLP (networks.py:37-57)."""

    def __init__(self, geometry, seed=1337):
        super().__init__()
        self.geometry = geometry
        gen = torch.Generator().manual_seed(seed)
        p = torch.empty(DENSITY_MLP_PARAMS + 2 * geometry.total)
        _xavier_uniform_(p[:2048], 64, 32, gen); _xavier_uniform_(p[2048:3072], 16, 64, gen)
        p[3072:] = (torch.rand(2 * geometry.total, generator=gen) * 2 - 1) * 1e-4  # tcnn: U(-1e-4, 1e-4)
        self.params = nn.Parameter(p)


class Encoding(nn.Module):
    """SH degree-4 direction encoder (networks.py:59-66): no parameters; kept for the state-dict key."""

    def __init__(self):
        super().__init__()
        self.params = nn.Parameter(torch.empty(0))

    def forward(self, u):
        """u = (d+1)/2 as the reference passes it (networks.py:145) -> (N,16) fp16."""
        d = (u.float() * 2 - 1).contiguous()
        out = torch.empty(d.shape[0], 16, dtype=torch.float16, device=d.device)
        call("arn_sh4", ptr(d), d.shape[0], ptr(out), stream())
        return out


class Network(nn.Module):
    """Parameter container of the colour MLP 32-64-64-3(16) (networks.py:68-78)."""

    def __init__(self, seed=1337):
        super().__init__()
        gen = torch.Generator().manual_seed(seed + 1)
        p = torch.empty(RGB_MLP_PARAMS)
        _xavier_uniform_(p[:2048], 64, 32, gen); _xavier_uniform_(p[2048:6144], 64, 64, gen)
        _xavier_uniform_(p[6144:], 16, 64, gen)
        self.params = nn.Parameter(p)


class _F16Cache:
    """fp16 working copy of an fp32 parameter, refreshed when the parameter's version changes (tiny-cuda-nn casts
    every forward; the fused Adam of trainer.py refreshes the copy itself and calls mark_fresh())."""

    def __init__(self):
        self.buf = None      # capacity >= numel (the sharded optimizer pads it to a whole number of shards)
        self.key = None
        self.capacity = 0

    def reserve(self, p, capacity):
        """(Re)allocate with room for `capacity` elements; the copy is re-cast on the next get()."""
        self.capacity = max(int(capacity), p.numel())
        self.buf = torch.empty(self.capacity, dtype=torch.float16, device=p.device)
        self.key = None

    def adopt(self, p, buf):
        """Use an externally owned fp16 buffer (peer-mapped memory of the multi-GPU exchange) as the working copy."""
        assert buf.dtype == torch.float16 and buf.numel() >= p.numel()
        self.buf, self.capacity, self.key = buf, buf.numel(), None

    def get(self, p):
        key = (p.data_ptr(), p._version, p.device)
        if self.buf is None or self.buf.numel() < p.numel() or self.buf.device != p.device:
            self.reserve(p, max(self.capacity, p.numel()))
        if key != self.key:
            src = p.detach()
            call("arn_cast_f32_to_f16", ptr(src), ptr(self.buf), src.numel(), stream())
            self.key = key
        return self.buf

    def mark_fresh(self, p):
        self.key = (p.data_ptr(), p._version, p.device)


class FieldState:
    """Everything a field evaluation needs besides the sample tensors."""

    def __init__(self, geometry, xyz_min, xyz_max, rgb_act='Sigmoid'):
        self.geometry = geometry
        self.set_box(xyz_min, xyz_max)
        self.rgb_act = 1 if rgb_act == 'Sigmoid' else 0
        self.loss_scale = DEFAULT_LOSS_SCALE
        self.cache_xyz, self.cache_rgb = _F16Cache(), _F16Cache()
        self.direct_grad = False  # trainer.py: accumulate straight into .grad buffers (no 46 MB temporaries)

    def set_box(self, xyz_min, xyz_max):
        self.mn = (C.c_float * 3)(*[float(v) for v in xyz_min]); self.mx = (C.c_float * 3)(*[float(v) for v in xyz_max])


def tile_rows(n):
    """Saved activations are stored as whole 128-row tiles (include/arnerf.h, arn_field_ws_t)."""
    return (n + 127) // 128 * 128


_WIMG = {}


def _scratch(device):
    """Inference calls share one scratch per device (nothing outlives the call); a training forward owns its own."""
    key = (device.type, device.index)
    if key not in _WIMG:
        _WIMG[key] = torch.empty(FIELD_SCRATCH_BYTES, dtype=torch.uint8, device=device)
    return _WIMG[key]


def _workspace(n, device, with_rgb, save=True, want_h=True):
    """save=False (no backward will follow: torch.no_grad / test-time rendering / occupancy refresh): only the feature
    tile image and, if asked for, h are allocated and the forward stores no activation."""
    e = lambda *s, dt=torch.float16: torch.empty(*s, dtype=dt, device=device)
    m = tile_rows(n)
    ws = dict(feat=e(m, 32))
    if want_h or save:
        ws["h"] = e(m, 16, dt=torch.float32)
    if not save:
        ws["wimg"] = _scratch(device)
        return ws
    ws.update(hid=e(m, 64), wimg=torch.empty(FIELD_SCRATCH_BYTES, dtype=torch.uint8, device=device))
    if with_rgb:
        ws.update(in32=e(m, 32), hid1=e(m, 64), hid2=e(m, 64))
    return ws


def _c_ws(ws):
    return FieldWs(ptr(ws["feat"]), ptr(ws.get("hid")), ptr(ws.get("h")), ptr(ws.get("in32")), ptr(ws.get("hid1")), ptr(ws.get("hid2")),
                   ptr(ws["wimg"]))


@torch.no_grad()
def field_inference(xyzs, dirs, params_xyz, params_rgb, state, impl="", want_h=False):
    """Forward only (no autograd graph, no saved activations): (sigmas, rgbs | None, h | None)."""
    xyzs = xyzs.contiguous().float()
    n, dev = xyzs.shape[0], xyzs.device
    with_rgb = dirs is not None
    p16x = state.cache_xyz.get(params_xyz)
    p16c = state.cache_rgb.get(params_rgb) if with_rgb else None
    save = impl != ""  # the CUDA-core cross-check kernels always write their activations
    ws = _workspace(n, dev, with_rgb, save=save, want_h=want_h)
    sigmas = torch.empty(n, dtype=torch.float32, device=dev)
    rgbs = torch.empty(n, 3, dtype=torch.float32, device=dev) if with_rgb else None
    if with_rgb:
        dirs = dirs.contiguous().float()
    call("arn_field_fw" + impl, ptr(xyzs), ptr(dirs), n, state.mn, state.mx, state.geometry.c_levels, ptr(p16x), ptr(p16c),
         state.rgb_act, _c_ws(ws), ptr(sigmas), ptr(rgbs), stream())
    return sigmas, rgbs, (ws["h"][:n] if "h" in ws else None)


class FieldFunction(torch.autograd.Function):
    """(xyzs, dirs | None, params_xyz, params_rgb | None) -> (sigmas (N) f32, rgbs (N,3) f32 | None, h (N,16) f32).

    Backward produces the hash-table + MLP gradients and, when xyzs requires grad, dL/dxyzs (render_surface_normal)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, xyzs, dirs, params_xyz, params_rgb, state, impl):
        xyzs = xyzs.contiguous()
        n, dev = xyzs.shape[0], xyzs.device
        with_rgb = dirs is not None
        p16x = state.cache_xyz.get(params_xyz)
        p16c = state.cache_rgb.get(params_rgb) if with_rgb else None
        ws = _workspace(n, dev, with_rgb)
        sigmas = torch.empty(n, dtype=torch.float32, device=dev)
        rgbs = torch.empty(n, 3, dtype=torch.float32, device=dev) if with_rgb else None
        if with_rgb:
            dirs = dirs.contiguous()
        g = state.geometry
        call("arn_field_fw" + impl, ptr(xyzs), ptr(dirs), n, state.mn, state.mx, g.c_levels, ptr(p16x), ptr(p16c),
             state.rgb_act, _c_ws(ws), ptr(sigmas), ptr(rgbs), stream())
        ctx.state, ctx.impl, ctx.ws, ctx.with_rgb = state, impl, ws, with_rgb
        ctx.p16 = (p16x, p16c)
        ctx.save_for_backward(xyzs, sigmas, rgbs, params_xyz, params_rgb)
        ctx.set_materialize_grads(False)
        h = ws["h"][:n]
        ctx.mark_non_differentiable(h)
        if with_rgb:
            return sigmas, rgbs, h
        return sigmas, None, h

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, d_sigmas, d_rgbs, d_h):
        xyzs, sigmas, rgbs, params_xyz, params_rgb = ctx.saved_tensors
        state, ws = ctx.state, ctx.ws
        n, dev = xyzs.shape[0], xyzs.device
        p16x, p16c = ctx.p16
        need_dx = ctx.needs_input_grad[0]
        need_px = ctx.needs_input_grad[2]
        need_pc = ctx.with_rgb and params_rgb is not None and ctx.needs_input_grad[3]
        d_sigmas = None if d_sigmas is None else d_sigmas.contiguous().float()
        d_rgbs = None if (d_rgbs is None or not ctx.with_rgb) else d_rgbs.contiguous().float()
        direct = state.direct_grad and need_px and params_xyz.grad is not None and \
            (not need_pc or params_rgb.grad is not None)
        if direct:
            gx = params_xyz.grad
            gc = params_rgb.grad if need_pc else None
        else:
            gx = torch.zeros_like(params_xyz, dtype=torch.float32)
            gc = torch.zeros(RGB_MLP_PARAMS, dtype=torch.float32, device=dev) if ctx.with_rgb else None
        dfeat = torch.empty(tile_rows(n), 32, dtype=torch.float32, device=dev)
        dx = torch.empty(n, 3, dtype=torch.float32, device=dev) if need_dx else None
        g = state.geometry
        call("arn_field_bw" + ctx.impl, ptr(xyzs), n, state.mn, state.mx, g.c_levels, ptr(p16x), ptr(p16c), state.rgb_act,
             _c_ws(ws), ptr(sigmas), ptr(rgbs), ptr(d_sigmas), ptr(d_rgbs), C.c_float(state.loss_scale), ptr(dfeat),
             ptr(gx), ptr(gc), ptr(dx), stream())
        if direct:
            return dx, None, None, None, None, None
        return dx, None, (gx if need_px else None), (gc if need_pc else None), None, None
