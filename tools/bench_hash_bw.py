"""Hash-grid backward on the W1 batch through the public entry point (arn_hash_encode_bw): device time per launch for the
"hash_bw_mode" (shortest run) x "hash_bw_blocks" (blocks per SM) tunables, with the L2 flushed between launches and with the
gradient buffer hot, and for level-group splits (arn_train_set_level_groups).  Scratch tool."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200 import _lib, vren
from ar_nerf_b200.networks import NGP
from ar_nerf_b200.workload import Workload

dev = torch.device("cuda:0")
w = Workload("W1"); model = NGP(0.5).to(dev); w.install(model); model.host_box()
ro, rd, tgt, noise = [t.to(dev) for t in w.train_batch(0)]
hits = vren.ray_aabb_near(ro, rd, [0, 0, 0], [0.5] * 3, 0.01)
out = vren.raymarching_train(ro, rd, hits[:, 0], model.density_bitfield, 1, 0.5, 0.0, noise, 128, 1024)
xyzs = out[1]; n = xyzs.shape[0]
st = model.field_state
dfeat = torch.randn(n, 32, device=dev)
tg = torch.zeros(model.geometry.total * 2, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("samples", n)


def run(iters=10, cold=True):
    ts = []
    for _ in range(iters):
        tg.zero_()
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("arn_hash_encode_bw", xyzs.data_ptr(), n, st.mn, st.mx, st.geometry.c_levels, None, dfeat.data_ptr(), tg.data_ptr(), None, _lib.stream())
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for blocks in (0, 2, 3):
    for mode in (1, 16):
        _lib.set_tunable("hash_bw_mode", mode); _lib.set_tunable("hash_bw_blocks", blocks)
        print(f"blocks/SM {blocks} min_run {mode:3d}: cold L2 {run():7.1f} us   warm {run(cold=False):7.1f} us")
_lib.set_tunable("hash_bw_mode", 1); _lib.set_tunable("hash_bw_blocks", 0)
for groups in ([0, 16], [0, 11, 16], [0, 11, 14, 16], [0, 8, 11, 13, 16], [0, 8, 16]):
    arr = (C.c_int * len(groups))(*groups)
    _lib.call("arn_train_set_level_groups", len(groups) - 1, arr, None)
    print(f"level groups {groups}: cold {run():7.1f} us  warm {run(cold=False):7.1f} us")
_lib.call("arn_train_set_level_groups", 0, None, None)
