"""Diagnostics: hash-grid backward variants per level range on the W1 batch (scratch tool)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200 import _lib
from ar_nerf_b200._lib import P, L, I, F, Levels
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.rendering import render
from ar_nerf_b200.workload import Workload
import ctypes as C

lib = _lib.lib()
lib.arn_dbg_hash_bw.argtypes = [P, L, P, P, Levels, P, P, I, I, I, P]
dev = torch.device("cuda:0")
w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model); model.host_box()
ro, rd, tgt, noise = [t.to(dev) for t in w.train_batch(0)]
from ar_nerf_b200 import vren
hits = vren.ray_aabb_near(ro, rd, [0, 0, 0], [0.5] * 3, 0.01)
out = vren.raymarching_train(ro, rd, hits[:, 0], model.density_bitfield, 1, 0.5, 0.0, noise, 128, 1024)
xyzs = out[1]; n = xyzs.shape[0]
print("samples", n)
st = model.field_state
dfeat = torch.randn(n, 32, device=dev)
ref = None
for name, l0, nl, mode in (("all per-sample", 0, 16, 0), ("all runs 8", 0, 16, 8), ("all runs 16", 0, 16, 16), ("all runs 32", 0, 16, 32), ("all runs 64", 0, 16, 64)):
    tg = torch.zeros(model.geometry.total * 2, device=dev)
    for it in range(3):
        tg.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.arn_dbg_hash_bw(xyzs.data_ptr(), n, st.mn, st.mx, st.geometry.c_levels, dfeat.data_ptr(), tg.data_ptr(), l0, nl, mode, _lib.stream())
        e1.record(); torch.cuda.synchronize()
        assert rc == 0
    print(f"{name:26s} {e0.elapsed_time(e1) * 1e3:8.1f} us")
    if name == "all per-sample": ref = tg.clone()
    else: print("   max diff vs per-sample", (tg - ref).abs().max().item(), "max", ref.abs().max().item())
