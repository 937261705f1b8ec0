"""The product's marching arithmetic (arn_march_core.h, shared by the CUDA kernels) compiled for the host and held
bit-exact to the oracle -- catches contraction/rounding mistakes before any GPU time is spent."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import ROOT, scene_hits

SO = os.path.join(ROOT, "tests", "_build", "host_march.so")


@pytest.fixture(scope="module")
def harness():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-o", SO,
                           os.path.join(ROOT, "tests", "host_march_harness.cpp")])
    return C.CDLL(SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_frexp_exponent(harness):
    harness.h_frexp_exponent.argtypes = [C.c_float]
    for x in [0.0, 1.0, 0.5, 0.49999997, 3.0, 1e-38, 1e-40, 1.4e-45, 16777216.0, -2.5, 0.00169]:
        assert harness.h_frexp_exponent(x) == math.frexp(np.float32(x))[1], x


@pytest.mark.parametrize("kind", ["W1", "W3"])
def test_train_march_bit_exact(harness, kind, w1, w3):
    w = w1 if kind == "W1" else w3
    ro, rd, _, noise = w.train_batch(1, 4096)
    ro, rd, noise = ro.numpy(), rd.numpy(), noise.numpy()
    ht = scene_hits(w, ro, rd); bits = w.bitfield.numpy()
    rays_a, xyzs, dirs, deltas, ts, counter = oracle.raymarching_train(ro, rd, ht, bits, w.cascades, w.scale, w.exp_step_factor, noise, 128, 1024)
    counts = np.zeros(len(ro), np.int32); rec = np.zeros((len(ro), 1024), np.float32)
    harness.h_march_train(len(ro), _p(ro), _p(rd), _p(ht), _p(bits), w.cascades, 128, C.c_float(w.scale), C.c_float(w.exp_step_factor),
                          _p(noise), 1024, _p(counts), _p(rec))
    assert np.array_equal(counts, rays_a[:, 2].astype(np.int32))
    got = np.concatenate([rec[r, :counts[r]] for r in range(len(ro))])
    assert np.array_equal(got.view(np.uint32), ts.view(np.uint32))
    # the warp-window procedure of march_train_count_warp_kernel, emulated lane by lane on the host
    for max_samples in (1024, 37):
        if max_samples != 1024:
            rays_a, xyzs, dirs, deltas, ts, counter = oracle.raymarching_train(ro, rd, ht, bits, w.cascades, w.scale, w.exp_step_factor, noise, 128, max_samples)
        counts = np.zeros(len(ro), np.int32); rec = np.zeros((len(ro), max_samples), np.float32)
        harness.h_march_train_window(len(ro), _p(ro), _p(rd), _p(ht), _p(bits), w.cascades, 128, C.c_float(w.scale), C.c_float(w.exp_step_factor),
                                     _p(noise), max_samples, _p(counts), _p(rec))
        assert np.array_equal(counts, rays_a[:, 2].astype(np.int32))
        got = np.concatenate([rec[r, :counts[r]] for r in range(len(ro))])
        assert np.array_equal(got.view(np.uint32), ts.view(np.uint32))


@pytest.mark.parametrize("kind", ["W1", "W3"])
def test_test_march_bit_exact(harness, kind, w1, w3):
    w = w1 if kind == "W1" else w3
    ro, rd, _, _ = w.train_batch(2, 2048)
    ro, rd = ro.numpy(), rd.numpy()
    ht = scene_hits(w, ro, rd); bits = w.bitfield.numpy()
    alive = np.arange(len(ro), dtype=np.int64)
    h1, h2 = ht.copy(), ht.copy()
    for S in (1, 4, 16, 64):
        x, d, dl, t, neff = oracle.raymarching_test(ro, rd, h1, alive, bits, w.cascades, w.scale, w.exp_step_factor, 128, 1024, S)
        ts = np.zeros((len(alive), S), np.float32); dls = np.zeros((len(alive), S), np.float32)
        xs = np.zeros((len(alive), S, 3), np.float32); ne = np.zeros(len(alive), np.int32)
        harness.h_march_test(len(alive), _p(ro), _p(rd), _p(h2), _p(alive), _p(bits), w.cascades, 128, C.c_float(w.scale),
                             C.c_float(w.exp_step_factor), S, 1024, _p(ts), _p(dls), _p(xs), _p(ne))
        assert np.array_equal(ne, neff) and np.array_equal(ts, t) and np.array_equal(dls, dl) and np.array_equal(xs, x)
        assert np.array_equal(h1, h2)
    # the warp-window form of the test march (march_test_warp_kernel), emulated lane by lane: same samples, same resume points
    h1, h3 = ht.copy(), ht.copy()
    for S in (9, 16, 64, 64, 33):
        _, _, dl, t, neff = oracle.raymarching_test(ro, rd, h1, alive, bits, w.cascades, w.scale, w.exp_step_factor, 128, 1024, S)
        ts = np.zeros((len(alive), S), np.float32); dls = np.zeros((len(alive), S), np.float32); ne = np.zeros(len(alive), np.int32)
        harness.h_march_test_window(len(alive), _p(ro), _p(rd), _p(h3), _p(alive), _p(bits), w.cascades, 128, C.c_float(w.scale),
                                    C.c_float(w.exp_step_factor), S, 1024, _p(ts), _p(dls), _p(ne))
        assert np.array_equal(ne, neff) and np.array_equal(ts, t) and np.array_equal(dls, dl)
        assert np.array_equal(h1, h3)


@pytest.mark.parametrize("kind", ["W1", "W3"])
def test_frame_marched_once_equals_sliced_test_march(harness, kind, w1, w3):
    """The claim behind arn_march_test_all / arn_render_test_step_pre: the test march is resumable and deterministic, so
    ONE march of a ray to its end yields exactly the samples the loop's iterations take slice by slice (oracle:
    raymarching_test called with the schedule's varying N_samples, rendering.py:189-203), with dt = calc_dt(t)."""
    w = w1 if kind == "W1" else w3
    ro, rd, _, _ = w.train_batch(3, 1024)
    ro, rd = ro.numpy(), rd.numpy()
    ht = scene_hits(w, ro, rd); bits = w.bitfield.numpy()
    R, stride = len(ro), 1088
    fast_ok = w.cascades == 1
    outs = []
    for fast in ([0, 1] if fast_ok else [0]):
        ts_all = np.zeros((stride, R), np.float32); dt_all = np.zeros((stride, R), np.float32); totals = np.zeros(R, np.int32)
        harness.h_march_test_all(R, _p(ro), _p(rd), _p(ht), _p(bits), w.cascades, 128, C.c_float(w.scale), C.c_float(w.exp_step_factor), 1024, stride,
                                 fast, _p(ts_all), _p(dt_all), _p(totals))
        outs.append((ts_all, dt_all, totals))
    if fast_ok:  # the compile-time shortcuts (one cascade, grid <= 256) change nothing
        assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))
    ts_all, dt_all, totals = outs[0]
    # the loop's slices from the oracle: every ray stays alive (no compositing here), N_samples varies like the schedule's
    h = ht.copy(); alive = np.arange(R, dtype=np.int64); cursor = np.zeros(R, np.int64)
    for S in (1, 2, 3, 7, 16, 64, 64, 5, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64):
        _, _, dl, t, neff = oracle.raymarching_test(ro, rd, h, alive, bits, w.cascades, w.scale, w.exp_step_factor, 128, 1024, S)
        assert np.array_equal(neff, np.minimum(S, totals - cursor))       # N_eff = min(S, total - cursor)
        for sl in range(S):
            m = sl < neff
            assert np.array_equal(t[m, sl], ts_all[(cursor + sl)[m], np.nonzero(m)[0]])
            assert np.array_equal(dl[m, sl], dt_all[(cursor + sl)[m], np.nonzero(m)[0]])
        cursor += neff
    assert (cursor <= totals).all() and cursor.sum() > 0
