"""Per-level-group device time of the hash-grid backward on the W1 batch: arn_train_set_level_groups with one event per
group, elapsed time between consecutive events.  Scratch tool."""
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
print("samples", n)


def per_group(groups, iters=7):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(groups))]
    for e in evs:
        e.record()
    arr = (C.c_int * len(groups))(*groups)
    handles = (C.c_void_p * (len(groups) - 1))(*[e.cuda_event for e in evs[1:]])
    _lib.call("arn_train_set_level_groups", len(groups) - 1, arr, handles)
    acc = [[] for _ in range(len(groups) - 1)]
    for _ in range(iters):
        tg.zero_()
        evs[0].record()
        _lib.call("arn_hash_encode_bw", xyzs.data_ptr(), n, st.mn, st.mx, st.geometry.c_levels, None, dfeat.data_ptr(), tg.data_ptr(), None, _lib.stream())
        torch.cuda.synchronize()
        for g in range(len(groups) - 1):
            acc[g].append(evs[g].elapsed_time(evs[g + 1]) * 1e3)
    _lib.call("arn_train_set_level_groups", 0, None, None)
    return [sorted(a)[len(a) // 2] for a in acc]


for mode in (1, 16):
    _lib.set_tunable("hash_bw_mode", mode)
    t = per_group(list(range(17)))
    print(f"min_run {mode}: per level us:", " ".join(f"{x:5.1f}" for x in t), " sum", round(sum(t), 1))
_lib.set_tunable("hash_bw_mode", 1)
for groups in ([0, 16], [0, 11, 16], [0, 8, 16], [0, 4, 8, 12, 16], [0, 11, 14, 16]):
    t = per_group(groups)
    print(f"groups {groups}: us per group:", " ".join(f"{x:5.1f}" for x in t), " sum", round(sum(t), 1))
