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
