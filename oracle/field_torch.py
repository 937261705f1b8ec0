"""Independent PyTorch restatement of the field and of the compositing (autograd does the backward) -- TEST
INFRASTRUCTURE ONLY.  Used to cross-check the hand-derived backward of oracle_field.c / oracle_vren.c and, through
them, the CUDA kernels.  Follows SURVEY Appendix A (tiny-cuda-nn semantics; parity unpinned) and
volumerendering.cu:5-44 / losses.cu.  fp16 rounding points are emulated with .half().float() through a
straight-through estimator so gradients flow as they do in the mixed-precision kernels."""
import torch

PRIMES = (1, 2654435761, 805459861)


class _RoundF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.half().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def f16(x):
    return _RoundF16.apply(x)


def grid_index(size, res, p):
    """p (..., 3) int64 grid coordinates -> index into the level (tiny-cuda-nn grid_index)."""
    stride, index = 1, torch.zeros_like(p[..., 0])
    for d in range(3):
        if stride <= size:
            index = index + p[..., d] * stride
            stride *= res
    if size < stride:
        index = ((p[..., 0] * PRIMES[0]) & 0xffffffff) ^ ((p[..., 1] * PRIMES[1]) & 0xffffffff) ^ ((p[..., 2] * PRIMES[2]) & 0xffffffff)
    return (index & 0xffffffff) % size


def hash_encode(x01, geo, table, dtype=torch.float32):
    """x01 (N,3) in [0,1]; table (total,2) (already fp16-rounded values); returns (N,32) un-rounded features."""
    outs = []
    for l in range(geo.n_levels):
        scale = float(geo.scale[l]); res = int(geo.res[l]); size = int(geo.size[l]); off = int(geo.offset[l])
        pos = x01.to(dtype) * scale + 0.5
        g = torch.floor(pos)
        w = pos - g
        g = g.long()
        acc = 0
        for c in range(8):
            wt = 1.0; p = []
            for d in range(3):
                if c & (1 << d):
                    wt = wt * w[:, d]; p.append(g[:, d] + 1)
                else:
                    wt = wt * (1 - w[:, d]); p.append(g[:, d])
            idx = off + grid_index(size, res, torch.stack(p, -1))
            acc = acc + wt[:, None] * table[idx].to(dtype)
        outs.append(acc)
    return torch.cat(outs, 1)


def sh4(dirs):
    d = dirs / torch.norm(dirs, dim=1, keepdim=True)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    return torch.stack([
        0.28209479177387814 * torch.ones_like(x), -0.48860251190291987 * y, 0.48860251190291987 * z,
        -0.48860251190291987 * x, 1.0925484305920792 * xy, -1.0925484305920792 * yz,
        0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * y * (-3.0 * x2 + y2),
        2.8906114426405538 * xy * z, 0.45704579946446572 * y * (1.0 - 5.0 * z2),
        0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)], 1)


class TruncExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        return g * torch.exp(ctx.saved_tensors[0].clamp(-15, 15))


def field(x01, dirs, geo, params_xyz, params_rgb, rgb_act=1, emulate_f16=True, dtype=torch.float32):
    """Returns sigmas (N), rgbs (N,3).  params_* are fp32 leaf tensors (gradients flow to them)."""
    r = f16 if emulate_f16 else (lambda t: t)
    px, pc = r(params_xyz.to(dtype)), r(params_rgb.to(dtype))
    W1, W2, table = px[:2048].view(64, 32), px[2048:3072].view(16, 64), px[3072:].view(-1, 2)
    feat = r(hash_encode(x01, geo, table, dtype))
    hid = r(torch.relu(feat @ W1.T))
    h = hid @ W2.T
    sigmas = TruncExp.apply(h[:, 0])
    inp = torch.cat([r(sh4(dirs.to(dtype))), r(h)], 1)
    C1, C2, C3 = pc[:2048].view(64, 32), pc[2048:6144].view(64, 64), pc[6144:].view(16, 64)
    a1 = r(torch.relu(inp @ C1.T)); a2 = r(torch.relu(a1 @ C2.T))
    o = (a2 @ C3.T)[:, :3]
    rgbs = torch.sigmoid(o) if rgb_act else o
    return sigmas, rgbs


def composite(sigmas, rgbs, deltas, ts, rays_a, T_threshold):
    """Differentiable front-to-back compositing with the reference's termination rule (volumerendering.cu:5-44)."""
    R = rays_a.shape[0]
    opacity, depth, rgb, ws = [], [], [], torch.zeros_like(sigmas)
    ws_parts = []
    for n in range(R):
        start, N = int(rays_a[n, 1]), int(rays_a[n, 2])
        s, d, c, t = sigmas[start:start + N], deltas[start:start + N], rgbs[start:start + N], ts[start:start + N]
        a = 1 - torch.exp(-s * d)
        T_after = torch.cumprod(1 - a, 0)
        T_before = torch.cat([torch.ones_like(T_after[:1]), T_after[:-1]])
        term = (T_after <= T_threshold).nonzero()
        keep = torch.ones(N, dtype=torch.bool)
        if len(term) > 0:
            keep[int(term[0]) + 1:] = False
        w = a * T_before * keep
        ws_parts.append(w)
        opacity.append(w.sum()); depth.append((w * t).sum()); rgb.append((w[:, None] * c).sum(0))
    order = torch.argsort(rays_a[:, 1], stable=True)
    ws = torch.cat([ws_parts[int(i)] for i in order]) if len(ws_parts) else ws
    out_o, out_d, out_c = torch.stack(opacity), torch.stack(depth), torch.stack(rgb)
    idx = rays_a[:, 0]
    O = torch.zeros_like(out_o).index_copy(0, idx, out_o); D = torch.zeros_like(out_d).index_copy(0, idx, out_d)
    C = torch.zeros_like(out_c).index_copy(0, idx, out_c)
    return O, D, C, ws


def distortion_loss(ws, deltas, ts, rays_a):
    """Direct O(N^2) definition of the mip-NeRF 360 distortion loss per ray (what losses.cu computes with scans)."""
    out = []
    for n in range(rays_a.shape[0]):
        start, N = int(rays_a[n, 1]), int(rays_a[n, 2])
        w, t, d = ws[start:start + N], ts[start:start + N], deltas[start:start + N]
        bi = (w[:, None] * w[None, :] * (t[:, None] - t[None, :]).abs()).sum()
        out.append(bi + (w * w * d).sum() / 3)
    return torch.stack(out)
