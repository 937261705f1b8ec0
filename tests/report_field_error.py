"""Measured error of the field kernels against the oracle, in the units the parity tests assert (calibration report for the constants of tests/test_gpu_parity.py; run on the GPU box: python tests/report_field_error.py):
  * forward: ambiguous-activation fraction, |h - h_oracle| against the a-priori rounding-interval bound, sigma, rgb;
  * backward: per-entry |grad - oracle| in units of eps32 * L1 (L1 = sum of |terms| of the entry, oracle.field_bw_l1).
Prints one line per tensor and implementation ("_simt" = CUDA cores in the oracle's order, "" = tcgen05)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle  # noqa: E402
import test_gpu_parity as tp  # noqa: E402

EPS32 = 2.0 ** -24


def main():
    for scale, n in ((0.5, 30000), (16.0, 8000)):
        for impl in ("_simt", ""):
            model, geo, x, x01, d, pxyz, prgb = tp._field_setup(scale, n, 5)
            model.field_impl = impl
            rng = np.random.default_rng(11)
            gs = (rng.standard_normal(n) * 1e-2).astype(np.float32); gc = (rng.standard_normal((n, 3)) * 1e-2).astype(np.float32)
            xt = tp.T(x).requires_grad_(True)
            sig, rgb = model(xt, tp.T(d))
            h = sig.grad_fn.ws["h"][:n]
            ((sig * tp.T(gs)).sum() + (rgb * tp.T(gc)).sum()).backward()
            ctx = oracle.field_fw(x01, d, geo, pxyz, prgb)
            bound, amb = tp.h_apriori_bound(ctx)
            eh = np.abs(tp.N(h).astype(np.float64) - ctx["h"])
            print(f"[scale {scale} impl '{impl}'] ambiguous hidden activations {amb.mean():.2e}; max |dh| {eh.max():.3e}; max |dh|/bound {np.max(eh / bound):.3f}; "
                  f"max rel sigma {np.max(np.abs(tp.N(sig) - ctx['sigma']) / ctx['sigma']):.3e}; max |drgb| {np.abs(tp.N(rgb) - ctx['rgb']).max():.3e}")
            o_gx, o_gc, o_dx, _ = oracle.field_bw(ctx, geo, gs, gc, loss_scale=128.0, want_dx=True)
            l1x, l1c = oracle.field_bw_l1(ctx, geo, gs, gc, loss_scale=128.0)
            for name, got, ref, l1 in (("colour MLP", tp.N(model.rgb_net.params.grad), o_gc, l1c),
                                       ("density MLP", tp.N(model.xyz_encoder.params.grad)[:3072], o_gx[:3072], l1x[:3072]),
                                       ("hash table", tp.N(model.xyz_encoder.params.grad)[3072:], o_gx[3072:], l1x[3072:])):
                err = np.abs(got.astype(np.float64) - ref)
                nz = l1 > 0
                r = err[nz] / (EPS32 * l1[nz])
                rel = err[nz] / np.maximum(np.abs(ref[nz]), 1e-300)
                print(f"    {name:12s} max err/(eps32 L1) {r.max():9.2f}  p99.9 {np.quantile(r, 0.999):8.2f}  | entries with rel err > 1e-4: {np.mean(rel > 1e-4):.3e}"
                      f"  max rel {rel.max():.2e}  | untouched entries nonzero: {int((got[~nz] != 0).sum())}  max|ref| {np.abs(ref).max():.3e}")


if __name__ == "__main__":
    main()
