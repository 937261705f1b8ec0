"""Compares the tensor-core field kernels with the simt cross-check path buffer by buffer (scratch diagnostics)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200 import _lib  # noqa: E402
from ar_nerf_b200.field import FieldFunction, _c_ws, _workspace  # noqa: E402
from ar_nerf_b200.networks import NGP  # noqa: E402
from ar_nerf_b200._lib import call, ptr, stream  # noqa: E402


def run(model, x, d, impl, n):
    st = model.field_state
    p16x = st.cache_xyz.get(model.xyz_encoder.params); p16c = st.cache_rgb.get(model.rgb_net.params)
    ws = _workspace(n, x.device, True)
    for v in ws.values():
        v.zero_()
    sig = torch.zeros(n, device=x.device); rgb = torch.zeros(n, 3, device=x.device)
    call("arn_field_fw" + impl, ptr(x), ptr(d), n, st.mn, st.mx, st.geometry.c_levels, ptr(p16x), ptr(p16c), 1, _c_ws(ws), ptr(sig), ptr(rgb), stream())
    torch.cuda.synchronize()
    return ws, sig, rgb, (p16x, p16c)


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = NGP(0.5).to(dev); model.host_box()
    with torch.no_grad():
        model.xyz_encoder.params[:3072].uniform_(-0.3, 0.3); model.xyz_encoder.params[3072:].uniform_(-0.5, 0.5)
        model.rgb_net.params.uniform_(-0.3, 0.3)
    for n in (128, 1000, 200000):
        x = (torch.rand(n, 3, device=dev) - 0.5); d = torch.randn(n, 3, device=dev)
        a_ws, a_sig, a_rgb, p16 = run(model, x, d, "_simt", n)
        b_ws, b_sig, b_rgb, _ = run(model, x, d, "_tc", n)
        print(f"n={n}")
        for k in ("feat", "hid", "h", "in32", "hid1", "hid2"):
            a, b = a_ws[k].float(), b_ws[k].float()
            print(f"  {k:5s} max|simt| {a.abs().max().item():.4f}  max diff {(a - b).abs().max().item():.3e}  mismatching rows {((a - b).abs().amax(1) > 1e-2 * a.abs().max()).sum().item()}")
        print(f"  sigma rel diff {((a_sig - b_sig).abs() / a_sig.abs().clamp_min(1e-6)).max().item():.3e}   rgb max diff {(a_rgb - b_rgb).abs().max().item():.3e}")
        # backward: simt vs tc on the simt forward's saved activations
        if "--bw" in sys.argv:
            st = model.field_state
            gs = torch.randn(n, device=dev) * 1e-2; gc = torch.randn(n, 3, device=dev) * 1e-2
            outs = {}
            for impl in ("_simt", "_tc"):
                gx = torch.zeros_like(model.xyz_encoder.params); gcw = torch.zeros(7168, device=dev); dfeat = torch.zeros(n, 32, device=dev)
                call("arn_field_bw" + impl, ptr(x), n, st.mn, st.mx, st.geometry.c_levels, ptr(p16[0]), ptr(p16[1]), 1, _c_ws(a_ws), ptr(a_sig), ptr(a_rgb),
                     ptr(gs), ptr(gc), 128.0, ptr(dfeat), ptr(gx), ptr(gcw), None, stream())
                torch.cuda.synchronize()
                outs[impl] = (gx, gcw, dfeat)
            for name, idx, sl in (("dWd", 0, slice(0, 3072)), ("dtable", 0, slice(3072, None)), ("dWc", 1, slice(None)), ("dfeat", 2, slice(None))):
                a, b = outs["_simt"][idx][sl], outs["_tc"][idx][sl]
                print(f"  {name:6s} max|simt| {a.abs().max().item():.4e}  max diff {(a - b).abs().max().item():.3e}")
            gcw = outs["_tc"][1]; a = outs["_simt"][1]
            for nm, s, e in (("C1", 0, 2048), ("C2", 2048, 6144), ("C3", 6144, 7168)):
                print(f"    dWc {nm}: max|simt| {a[s:e].abs().max().item():.3e} diff {(a[s:e] - gcw[s:e]).abs().max().item():.3e}")
            gx_t, gx_s = outs["_tc"][0], outs["_simt"][0]
            for nm, s, e in (("D1", 0, 2048), ("D2", 2048, 3072)):
                print(f"    dWd {nm}: max|simt| {gx_s[s:e].abs().max().item():.3e} diff {(gx_s[s:e] - gx_t[s:e]).abs().max().item():.3e}")


if __name__ == "__main__":
    main()
