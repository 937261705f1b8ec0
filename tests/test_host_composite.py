"""Lane-by-lane host emulation (numpy float32, no GPU) of the training compositing kernels' turn structure
(ar_nerf_b200/csrc/arn_vren.cu: composite_train_fw_kernel / composite_bw_ray -- a warp takes kSPL * 32 consecutive samples of
its ray per turn, lane l the samples kSPL*l .. kSPL*l + kSPL-1, scanned inside the lane, the lanes' totals by one
Hillis-Steele warp scan per quantity) against the oracle's sequential restatement of volumerendering.cu:5-44,86-150:
the regrouped products and sums stay within a few float32 roundings of the sequential ones, the early-termination index and
the zeroing behind it are the reference's."""
import numpy as np
import pytest

import oracle

F = np.float32
K_SPL = 2  # arn_vren.cu: kSPL


def _warp_incl(vals, op):
    """Hillis-Steele inclusive scan over 32 lanes in float32 (warp_incl_sum / warp_incl_prod)."""
    v = vals.astype(F).copy()
    o = 1
    while o < 32:
        up = np.empty_like(v); up[o:] = v[:-o]
        nv = v.copy()
        nv[o:] = op(v[o:], up[o:]).astype(F)
        v = nv
        o <<= 1
    return v


def _butterfly_sum(vals):
    """warp_sum: xor-butterfly; every lane ends with the same float32 total."""
    v = vals.astype(F).copy()
    o = 16
    while o > 0:
        v = (v + v[np.arange(32) ^ o]).astype(F)
        o >>= 1
    return v[0]


def _alpha(sigma, delta):
    return (F(1.0) - np.exp(-(sigma * delta).astype(F)).astype(F)).astype(F)


def emulate_fw(sigmas, rgbs, deltas, ts, rays_a, thr):
    R = len(rays_a)
    total = np.zeros(R, np.int64); opacity = np.zeros(R, F); depth = np.zeros(R, F); rgb = np.zeros((R, 3), F)
    ws = np.zeros(len(sigmas), F)
    chunk = 32 * K_SPL
    for ray_idx, start, N in rays_a:
        T = F(1.0); acc = np.zeros(5, F); samples = N; done = False; base = 0
        while base < N and not done:
            idx = base + np.arange(chunk)                     # lane l owns idx[K_SPL*l : K_SPL*l + K_SPL]
            inside = idx < N
            q = start + np.minimum(idx, max(N - 1, 0))
            al = np.where(inside, _alpha(sigmas[q], deltas[q]), F(0)).astype(F).reshape(32, K_SPL)
            om = (F(1.0) - al).astype(F)
            pp = np.empty_like(om); pp[:, 0] = om[:, 0]
            for k in range(1, K_SPL):
                pp[:, k] = (pp[:, k - 1] * om[:, k]).astype(F)
            incl = _warp_incl(pp[:, -1], np.multiply)
            excl = np.concatenate([[F(1.0)], incl[:-1]]).astype(F)
            Tl = (T * excl).astype(F)
            Ta = (Tl[:, None] * pp).astype(F)
            Tb = np.empty_like(Ta); Tb[:, 0] = Tl; Tb[:, 1:] = (Tl[:, None] * pp[:, :-1]).astype(F)
            term = inside.reshape(32, K_SPL) & (Ta <= F(thr))
            last = chunk - 1
            if term.any():
                last = int(np.argmax(term.reshape(-1)))       # first terminating sample, which still contributes
            use = inside & (np.arange(chunk) <= last)
            w = np.where(use.reshape(32, K_SPL), (al * Tb).astype(F), F(0)).astype(F)
            ws[start + idx[inside]] = w.reshape(-1)[inside]
            vals = [rgbs[q, 0], rgbs[q, 1], rgbs[q, 2], ts[q], np.ones(chunk, F)]
            for j, c in enumerate(vals):
                prod = (w * c.astype(F).reshape(32, K_SPL)).astype(F)
                lane = np.zeros(32, F)
                for k in range(K_SPL):
                    lane = (lane + prod[:, k]).astype(F)
                acc[j] = F(acc[j] + _butterfly_sum(lane))
            if term.any():
                done = True; samples = base + last
            T = F(Tl[31] * pp[31, -1])
            base += chunk
        rgb[ray_idx] = acc[:3]; depth[ray_idx] = acc[3]; opacity[ray_idx] = acc[4]; total[ray_idx] = samples
    return total, opacity, depth, rgb, ws


def _rays(rng, counts, sigma_max):
    counts = np.asarray(counts, np.int64)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    rays_a = np.stack([rng.permutation(len(counts)), starts, counts], 1).astype(np.int64)
    n = int(counts.sum())
    sigmas = (rng.random(n) * sigma_max).astype(F)
    rgbs = rng.random((n, 3)).astype(F)
    deltas = np.full(n, 1.7320508 / 1024, F)
    ts = np.concatenate([0.3 + np.cumsum(deltas[s:s + c]) for s, c in zip(starts, counts)]).astype(F) if n else np.zeros(0, F)
    return sigmas, rgbs, deltas, ts, rays_a


@pytest.mark.parametrize("sigma_max,thr", [(40.0, 1e-4), (4000.0, 1e-4), (4000.0, 1e-2)])
def test_turn_structure_matches_sequential_compositing(sigma_max, thr):
    rng = np.random.default_rng(5)
    counts = [0, 1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 300, 437, 0, 7]
    sigmas, rgbs, deltas, ts, rays_a = _rays(rng, counts, sigma_max)
    want = oracle.composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, thr)
    got = emulate_fw(sigmas, rgbs, deltas, ts, rays_a, thr)
    # a ray may cross T <= thr one sample earlier or later when T lands within rounding of thr: none of these does
    assert np.array_equal(got[0], want[0]), "termination index"
    if sigma_max > 1000:
        assert (want[0] < np.asarray(counts)[np.argsort(rays_a[:, 0])]).any(), "the case must exercise early termination"
    for name, g, w_ in zip(("opacity", "depth", "rgb", "ws"), got[1:], want[1:]):
        tol = 2e-6 * np.abs(w_) + 4 * 2.0 ** -24   # the weights carry alpha = 1 - exp(..): an absolute floor of a few ulp(1)
        assert np.all(np.abs(g.astype(np.float64) - w_) <= tol), name
    # behind the terminating sample the weights are exactly zero, as the reference's zero-initialised output
    for ray_idx, start, N in rays_a:
        used = int(want[0][ray_idx])
        assert not got[4][start + used + 1:start + N].any()


def emulate_bw(g_op, g_dep, g_rgb, sigmas, rgbs, deltas, ts, rays_a, opacity, depth, rgb, thr):
    """composite_bw_ray without a dL_dws input (the fused training step's form): in-lane inclusive sums, one warp scan of the
    lanes' totals per quantity, running carries across turns."""
    dsig = np.zeros(len(sigmas), F); drgbs = np.zeros((len(sigmas), 3), F)
    chunk = 32 * K_SPL
    for ray_idx, start, N in rays_a:
        R_, G_, B_ = (F(v) for v in rgb[ray_idx]); O_, D_ = F(opacity[ray_idx]), F(depth[ray_idx])
        gR, gG, gB = (F(v) for v in g_rgb[ray_idx]); gO, gD = F(g_op[ray_idx]), F(g_dep[ray_idx])
        T = F(1.0); carry = np.zeros(4, F); done = False; base = 0
        while base < N and not done:
            idx = base + np.arange(chunk); inside = idx < N
            q = start + np.minimum(idx, max(N - 1, 0))
            shape = (32, K_SPL)
            dl = np.where(inside, deltas[q], F(0)).astype(F).reshape(shape)
            al = np.where(inside, _alpha(sigmas[q], deltas[q]), F(0)).astype(F).reshape(shape)
            om = (F(1.0) - al).astype(F)
            pp = np.empty_like(om); pp[:, 0] = om[:, 0]
            for k in range(1, K_SPL):
                pp[:, k] = (pp[:, k - 1] * om[:, k]).astype(F)
            incl = _warp_incl(pp[:, -1], np.multiply)
            Tl = (T * np.concatenate([[F(1.0)], incl[:-1]]).astype(F)).astype(F)
            Ta = (Tl[:, None] * pp).astype(F)
            Tb = np.empty_like(Ta); Tb[:, 0] = Tl; Tb[:, 1:] = (Tl[:, None] * pp[:, :-1]).astype(F)
            term = inside.reshape(shape) & (Ta <= F(thr))
            last = int(np.argmax(term.reshape(-1))) if term.any() else chunk - 1
            use = (inside & (np.arange(chunk) <= last)).reshape(shape)
            w = np.where(use, (al * Tb).astype(F), F(0)).astype(F)
            cols = [np.where(inside, c[q], F(0)).astype(F).reshape(shape) for c in (rgbs[:, 0], rgbs[:, 1], rgbs[:, 2], ts)]
            pre = []                                                   # per-sample inclusive prefix of w*c over the whole ray
            for j, c in enumerate(cols):
                loc = np.cumsum((w * c).astype(F), axis=1, dtype=F)    # in-lane, sequential in float32
                tot = _warp_incl(loc[:, -1], np.add)
                excl = (carry[j] + (tot - loc[:, -1]).astype(F)).astype(F)
                pre.append((excl[:, None] + loc).astype(F))
                carry[j] = F(carry[j] + tot[31])
            pr, pg, pb, pd = pre
            cr, cg, cb, ct = cols
            o_s = (dl * (gR * (cr * Ta - (R_ - pr)) + gG * (cg * Ta - (G_ - pg)) + gB * (cb * Ta - (B_ - pb))
                         + gO * (F(1.0) - O_) + gD * (ct * Ta - (D_ - pd)))).astype(F)
            sel = inside.reshape(shape)
            dsig[start + idx[inside]] = np.where(use, o_s, F(0))[sel]
            for c_, g_ in enumerate((gR, gG, gB)):
                drgbs[start + idx[inside], c_] = np.where(use, (g_ * w).astype(F), F(0))[sel]
            if term.any():
                done = True
            T = F(Ta[31, -1])
            base += chunk
    return dsig, drgbs


@pytest.mark.parametrize("sigma_max,thr", [(40.0, 1e-4), (4000.0, 1e-2)])
def test_turn_structure_matches_sequential_backward(sigma_max, thr):
    rng = np.random.default_rng(9)
    counts = [0, 1, 33, 64, 65, 128, 200, 437, 5]
    sigmas, rgbs, deltas, ts, rays_a = _rays(rng, counts, sigma_max)
    _, opacity, depth, rgb, ws = oracle.composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, thr)
    R = len(rays_a)
    g_op, g_dep, g_rgb = rng.standard_normal(R).astype(F), rng.standard_normal(R).astype(F), rng.standard_normal((R, 3)).astype(F)
    want_s, want_c = oracle.composite_train_bw(g_op, g_dep, g_rgb, np.zeros(len(sigmas), F), sigmas, rgbs, ws, deltas, ts, rays_a,
                                               opacity, depth, rgb, thr)
    got_s, got_c = emulate_bw(g_op, g_dep, g_rgb, sigmas, rgbs, deltas, ts, rays_a, opacity, depth, rgb, thr)
    assert np.all(np.abs(got_c.astype(np.float64) - want_c) <= 2e-6 * np.abs(want_c) + 1e-6), "dL_drgbs"
    # dL_dsigma is a difference of ray-sized sums scaled by delta: judged against the size of its terms, not of the result
    scale = deltas * (np.abs(g_rgb).sum(1).max() * 4 + np.abs(g_op).max() + np.abs(g_dep).max() * ts.max(initial=1.0))
    assert np.all(np.abs(got_s.astype(np.float64) - want_s) <= 1e-4 * np.abs(want_s) + 4e-6 * scale), "dL_dsigmas"
