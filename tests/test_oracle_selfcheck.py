"""Self-checks of the CPU oracle (the reference ships no tests or golden vectors for this path).
They hold the oracle to closed forms, numpy, an independent autograd restatement and structural invariants.
The pin against the REAL reference kernels is tests/test_oracle_golden.py."""
import numpy as np
import pytest
import torch

import oracle
from oracle import field_torch
from conftest import scene_hits


def test_morton_roundtrip_and_known_values():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 1024, (20000, 3)).astype(np.int32)
    idx = oracle.morton3D(c)
    assert (oracle.morton3D_invert(idx) == c).all()
    assert oracle.morton3D(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [3, 3, 3], [127, 127, 127]], np.int32)).tolist() == [1, 2, 4, 63, 2097151]


def test_packbits_matches_numpy_little():
    rng = np.random.default_rng(1)
    g = rng.random(128 * 64).astype(np.float32) * 10
    g[::7] = 5.0  # strict '>' at the threshold
    bits = np.zeros(g.size // 8, np.uint8)
    oracle.packbits(g, 5.0, bits)
    assert (bits == np.packbits(g > 5.0, bitorder='little')).all()


@pytest.mark.parametrize("kind", ["W1", "W3"])
def test_march_invariants(kind, w1, w3):
    w = w1 if kind == "W1" else w3
    ro, rd, _, noise = w.train_batch(0, 2048)
    ro, rd, noise = ro.numpy(), rd.numpy(), noise.numpy()
    ht = scene_hits(w, ro, rd)
    bits = w.bitfield.numpy()
    rays_a, xyzs, dirs, deltas, ts, counter = oracle.raymarching_train(ro, rd, ht, bits, w.cascades, w.scale, w.exp_step_factor, noise, 128, 1024)
    assert counter[0] == rays_a[:, 2].sum() == len(ts) and counter[0] > 0
    assert (rays_a[:, 0] == np.arange(len(ro))).all()
    assert (rays_a[:, 1] == np.concatenate([[0], np.cumsum(rays_a[:, 2])[:-1]])).all()
    assert rays_a[:, 2].max() <= 1024
    # rays that miss the box march nothing
    assert (rays_a[ht[:, 0] < 0, 2] == 0).all()
    # per ray: ts strictly increase by at least the previous delta, samples lie on the ray, inside [t1, t2)
    for r in np.nonzero(rays_a[:, 2])[0][:200]:
        s, n = rays_a[r, 1], rays_a[r, 2]
        t, d = ts[s:s + n], deltas[s:s + n]
        assert (np.diff(t) >= d[:-1] * 0.999).all()
        assert t[0] >= ht[r, 0] and t[-1] < ht[r, 1]
        np.testing.assert_allclose(xyzs[s:s + n], ro[r] + t[:, None] * rd[r], rtol=0, atol=1e-6 * max(1, w.scale))
        assert (dirs[s:s + n] == rd[r]).all()
    # every sample sits in an occupied cell of its cascade (cascade 0 check for W1)
    if kind == "W1":
        n = np.clip((0.5 * (xyzs / 0.5 + 1) * 128), 0, 127).astype(np.int32)
        idx = oracle.morton3D(n).astype(np.int64)
        occ = (bits[idx // 8] >> (idx % 8)) & 1
        assert occ.mean() > 0.999  # float rounding of the re-derived cell can differ on a boundary


def test_march_test_resumes_to_train_samples(w1):
    """Chunked test-time marching (raymarching.cu:335-404) visits exactly the samples of a noise-free train march."""
    w = w1
    ro, rd, _, _ = w.train_batch(3, 1024)
    ro, rd = ro.numpy(), rd.numpy()
    ht = scene_hits(w, ro, rd)
    bits = w.bitfield.numpy()
    rays_a, xyzs, _, deltas, ts, _ = oracle.raymarching_train(ro, rd, ht, bits, 1, 0.5, 0.0, np.zeros(len(ro), np.float32), 128, 1024)
    hits = ht.copy()
    alive = np.arange(len(ro), dtype=np.int64)
    got = [[] for _ in range(len(ro))]
    for S in (1, 2, 4, 7, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64):
        if len(alive) == 0:
            break
        x, d, dl, t, neff = oracle.raymarching_test(ro, rd, hits, alive, bits, 1, 0.5, 0.0, 128, 1024, S)
        for k, r in enumerate(alive):
            got[r] += list(t[k, :neff[k]])
            assert (t[k, neff[k]:] == 0).all() and (x[k, neff[k]:] == 0).all()
        alive = alive[neff > 0]
    for r in range(len(ro)):
        s, n = rays_a[r, 1], rays_a[r, 2]
        # scale=0.5, cascades=1: the test kernel's calc_dt(.., cascades) quirk only moves the (unused) upper clamp
        assert np.array_equal(np.array(got[r], np.float32), ts[s:s + n])


def test_aabb_semantics():
    o = np.array([[0, 0, -2], [0, 0, -2], [0, 0, 0], [2, 2, 2], [0, 0, -2]], np.float32)
    d = np.array([[0, 0, 1], [0, 1, 0], [0, 0, 1], [1, 0, 0], [1e-3, 0, 1]], np.float32)
    cnt, ht, hi = oracle.ray_aabb_intersect(o, d, np.zeros((1, 3), np.float32), np.full((1, 3), 0.5, np.float32), 1)
    assert cnt.tolist() == [1, 0, 1, 0, 1]
    np.testing.assert_allclose(ht[0, 0], [1.5, 2.5])        # axis-parallel ray: 0*inf NaNs dropped by fminf/fmaxf
    assert (ht[1, 0] == -1).all() and (ht[3, 0] == -1).all() and hi[1, 0] == -1
    np.testing.assert_allclose(ht[2, 0], [0.0, 0.5])        # origin inside: t1 = max(t1, 0)
    # several voxels, max_hits 3: unfilled (-1) slots sort first, then near to far
    centers = np.array([[0, 0, 0], [0, 0, 2]], np.float32); half = np.full((2, 3), 0.5, np.float32)
    cnt, ht, hi = oracle.ray_aabb_intersect(o[:1], d[:1], centers, half, 3)
    assert cnt[0] == 2 and hi[0].tolist() == [-1, 0, 1] and ht[0, 0, 0] == -1 and ht[0, 1, 0] == 1.5 and ht[0, 2, 0] == 3.5


def _random_rays_a(rng, R, max_n):
    n = rng.integers(0, max_n, R)
    n[rng.random(R) < 0.2] = 0
    start = np.concatenate([[0], np.cumsum(n)[:-1]])
    return np.stack([np.arange(R), start, n], 1).astype(np.int64), int(n.sum())


def test_composite_constant_sigma_closed_form():
    N, sigma, delta = 40, 3.0, 0.05
    rays_a = np.array([[0, 0, N]], np.int64)
    ts = (np.arange(N) * delta).astype(np.float32)
    total, opacity, depth, rgb, ws = oracle.composite_train_fw(np.full(N, sigma, np.float32), np.full((N, 3), 0.25, np.float32),
                                                               np.full(N, delta, np.float32), ts, rays_a, 0.0)
    a = 1 - np.exp(-sigma * delta)
    w = a * (1 - a) ** np.arange(N)
    np.testing.assert_allclose(ws, w, rtol=2e-5)
    np.testing.assert_allclose(opacity[0], 1 - (1 - a) ** N, rtol=2e-5)
    np.testing.assert_allclose(rgb[0], 0.25 * (1 - (1 - a) ** N), rtol=2e-5)
    assert total[0] == N


def test_composite_early_termination_quirks():
    """Q4/Q5: terminated rays count index-of-last-sample; ws past the terminating sample is exactly 0."""
    N = 50
    rays_a = np.array([[0, 0, N]], np.int64)
    total, opacity, depth, rgb, ws = oracle.composite_train_fw(np.full(N, 100.0, np.float32), np.ones((N, 3), np.float32),
                                                               np.full(N, 0.05, np.float32), np.arange(N, dtype=np.float32), rays_a, 1e-4)
    k = int(total[0])
    assert 0 < k < N and ws[k] > 0 and (ws[k + 1:] == 0).all()


def test_composite_fw_bw_vs_autograd():
    rng = np.random.default_rng(5)
    rays_a, N = _random_rays_a(rng, 64, 60)
    sig = (rng.random(N) * 30).astype(np.float32); rgbs = rng.random((N, 3)).astype(np.float32)
    deltas = (rng.random(N) * 0.01 + 0.002).astype(np.float32); ts = np.cumsum(deltas).astype(np.float32)
    thr = 1e-2
    total, opacity, depth, rgb, ws = oracle.composite_train_fw(sig, rgbs, deltas, ts, rays_a, thr)
    ts_, rg_ = torch.tensor(sig, dtype=torch.float64, requires_grad=True), torch.tensor(rgbs, dtype=torch.float64, requires_grad=True)
    O, D, C, W = field_torch.composite(ts_, rg_, torch.tensor(deltas, dtype=torch.float64), torch.tensor(ts, dtype=torch.float64),
                                       torch.tensor(rays_a), thr)
    np.testing.assert_allclose(opacity, O.detach().numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(depth, D.detach().numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(rgb, C.detach().numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ws, W.detach().numpy(), rtol=1e-4, atol=1e-7)
    gO, gD, gC, gW = (rng.standard_normal(64), rng.standard_normal(64), rng.standard_normal((64, 3)), rng.standard_normal(N))
    (O * torch.tensor(gO) + D * torch.tensor(gD)).sum().add((C * torch.tensor(gC)).sum()).add((W * torch.tensor(gW)).sum()).backward()
    dsig, drgbs = oracle.composite_train_bw(gO, gD, gC, gW, sig, rgbs, ws, deltas, ts, rays_a, opacity, depth, rgb, thr)
    # the reference's analytic backward treats the termination as non-differentiable, like the masked autograd form
    np.testing.assert_allclose(drgbs, rg_.grad.numpy(), rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(dsig, ts_.grad.numpy(), rtol=2e-3, atol=2e-5)


def test_distortion_loss_vs_definition_and_autograd():
    rng = np.random.default_rng(6)
    rays_a, N = _random_rays_a(rng, 32, 40)
    ws = rng.random(N).astype(np.float32) * 0.1
    deltas = (rng.random(N) * 0.01 + 0.002).astype(np.float32); ts = np.cumsum(deltas).astype(np.float32)
    loss, wsi, wtsi = oracle.distortion_loss_fw(ws, deltas, ts, rays_a)
    w_ = torch.tensor(ws, dtype=torch.float64, requires_grad=True)
    ref = field_torch.distortion_loss(w_, torch.tensor(deltas, dtype=torch.float64), torch.tensor(ts, dtype=torch.float64), torch.tensor(rays_a))
    np.testing.assert_allclose(loss, ref.detach().numpy(), rtol=2e-3, atol=1e-7)
    g = rng.standard_normal(32).astype(np.float32)
    (ref * torch.tensor(g, dtype=torch.float64)).sum().backward()
    dws = oracle.distortion_loss_bw(g, wsi, wtsi, ws, deltas, ts, rays_a)
    np.testing.assert_allclose(dws, w_.grad.numpy(), rtol=5e-3, atol=2e-6)


def test_hash_geometry_table():
    """SURVEY Appendix A.2 (float32 host arithmetic: level 5 of scale 0.5 is 65^3, not 64^3)."""
    g = oracle.HashGeometry(per_level_scale=float(np.float32(np.exp(np.log(2048 * 0.5 / 16) / 15))))
    assert g.res.tolist() == [16, 22, 28, 37, 49, 65, 85, 112, 148, 195, 257, 338, 446, 589, 777, 1025]
    assert g.size[:6].tolist() == [4096, 10648, 21952, 50656, 117656, 274632] and (g.size[6:] == 2 ** 19).all()
    assert g.total == 5722520
    g = oracle.HashGeometry(per_level_scale=float(np.float32(np.exp(np.log(2048 * 16 / 16) / 15))))
    assert g.res[:4].tolist() == [16, 27, 45, 74] and g.total == 6811592


def _field_inputs(n=2000, seed=7, table_amp=0.5):
    rng = np.random.default_rng(seed)
    geo = oracle.HashGeometry(per_level_scale=float(np.float32(np.exp(np.log(2048 * 0.5 / 16) / 15))))
    x01 = rng.random((n, 3)).astype(np.float32)
    x01[:8] = [[0, 0, 0], [1, 1, 1], [0, 1, 0], [1, 0, 0], [0.5, 0.5, 0.5], [1, 1, 0], [0, 0, 1], [0.999999, 0.5, 0]]
    dirs = rng.standard_normal((n, 3)).astype(np.float32)
    pxyz = np.concatenate([(rng.random(3072) * 2 - 1) * 0.3, (rng.random(2 * geo.total) * 2 - 1) * table_amp]).astype(np.float32)
    prgb = ((rng.random(7168) * 2 - 1) * 0.3).astype(np.float32)
    return geo, x01, dirs, pxyz, prgb


def test_field_forward_vs_torch_restatement():
    geo, x01, dirs, pxyz, prgb = _field_inputs()
    ctx = oracle.field_fw(x01, dirs, geo, pxyz, prgb)
    with torch.no_grad():
        px, pc = torch.tensor(pxyz), torch.tensor(prgb)
        table = px[3072:].half().float().view(-1, 2)
        feat = field_torch.hash_encode(torch.tensor(x01), geo, table)
        sig, rgb = field_torch.field(torch.tensor(x01), torch.tensor(dirs), geo, px, pc)
    np.testing.assert_allclose(ctx["feat"].astype(np.float32), feat.numpy(), rtol=2e-3, atol=2e-4)  # one fp16 ulp
    np.testing.assert_allclose(ctx["sigma"], sig.numpy(), rtol=2e-2)
    np.testing.assert_allclose(ctx["rgb"], rgb.numpy(), atol=5e-3)


def test_field_backward_vs_autograd():
    geo, x01, dirs, pxyz, prgb = _field_inputs(n=1500, seed=8)
    rng = np.random.default_rng(9)
    ctx = oracle.field_fw(x01, dirs, geo, pxyz, prgb)
    gs = (rng.standard_normal(len(x01)) * 1e-2).astype(np.float32); gc = (rng.standard_normal((len(x01), 3)) * 1e-2).astype(np.float32)
    gx, gcw, dx, dfeat = oracle.field_bw(ctx, geo, gs, gc, loss_scale=128.0, want_dx=True)
    px = torch.tensor(pxyz, requires_grad=True); pc = torch.tensor(prgb, requires_grad=True)
    xt = torch.tensor(x01, requires_grad=True)
    sig, rgb = field_torch.field(xt, torch.tensor(dirs), geo, px, pc)
    ((sig * torch.tensor(gs)).sum() + (rgb * torch.tensor(gc)).sum()).backward()

    def close(a, b, tol):
        a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
        assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30), (np.abs(a - b).max(), np.abs(b).max())
    close(gcw, pc.grad.numpy(), 2e-2)                       # colour MLP weights
    close(gx[:3072], px.grad.numpy()[:3072], 2e-2)          # density MLP weights
    close(gx[3072:], px.grad.numpy()[3072:], 2e-2)          # hash table
    # dL/dx through the trilinear weights only (the oracle and tcnn differentiate the interpolation, not floor)
    close(dx, xt.grad.numpy(), 5e-2)


def test_field_bw_l1_dominates_the_gradient():
    """oracle.field_bw_l1 (all-paths sum of |products| per parameter, the yardstick of assert_sum in the GPU parity tests):
    entry by entry >= |gradient| (triangle inequality; the gradient's fp16-rounded factors may exceed the unrounded
    magnitudes by 2^-11 per rounding), zero exactly where no sample contributes, and equal to |gradient| for the one weight
    block whose products all share a sign when the upstream gradient does (the output layer, non-negative activations)."""
    geo, x01, dirs, pxyz, prgb = _field_inputs(n=1500, seed=8)
    rng = np.random.default_rng(10)
    ctx = oracle.field_fw(x01, dirs, geo, pxyz, prgb)
    gs = (rng.standard_normal(len(x01)) * 1e-2).astype(np.float32); gc = np.abs(rng.standard_normal((len(x01), 3)) * 1e-2).astype(np.float32)
    gx, gcw, _, _ = oracle.field_bw(ctx, geo, gs, gc, loss_scale=128.0)
    l1x, l1c = oracle.field_bw_l1(ctx, geo, gs, gc, loss_scale=128.0)
    assert (l1x >= np.abs(gx) * (1 - 2e-3)).all() and (l1c >= np.abs(gcw) * (1 - 2e-3)).all()
    assert np.array_equal(l1x[3072:] == 0, gx[3072:] == 0) or not (gx[3072:][l1x[3072:] == 0] != 0).any()
    out_rows = slice(6144, 6144 + 3 * 64)  # dW3 rows of the three colour outputs: g3 >= 0 (positive upstream, sigmoid' > 0), hid2 >= 0
    np.testing.assert_allclose(l1c[out_rows], np.abs(gcw[out_rows]), rtol=1e-3)
    assert l1c.sum() > np.abs(gcw).sum() * 1.5  # cancellation is real elsewhere
