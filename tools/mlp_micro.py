"""Device time of the field's MLP kernels alone, by sample count and mode (saved activations or not).  Scratch diagnostics."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_nerf_b200 import _lib  # noqa: E402
from ar_nerf_b200.field import FieldFunction, field_inference  # noqa: E402
from ar_nerf_b200.networks import NGP  # noqa: E402
from ar_nerf_b200.workload import Workload  # noqa: E402


def prof(fn, reps=10):
    for _ in range(3):
        fn()
    _lib.profile_enable(True)
    for _ in range(reps):
        fn()
    s = _lib.profile_report(); _lib.profile_enable(False)
    return {k: ms / reps * 1e3 for k, (n, ms) in s.items()}


def main():
    dev = torch.device("cuda:0")
    w = Workload("W1")
    model = NGP(0.5).to(dev)
    w.install(model)
    st = model.field_state
    g = torch.Generator(device=dev).manual_seed(1)
    for n in (61705, 123410, 246821, 493642, 987284, 2560000):
        x = (torch.rand(n, 3, device=dev, generator=g) - 0.5) * 0.9
        d = torch.nn.functional.normalize(torch.randn(n, 3, device=dev, generator=g), dim=1)
        px, pc = model.xyz_encoder.params, model.rgb_net.params
        inf = prof(lambda: field_inference(x, d, px, pc, st))
        ds, dr = torch.randn(n, device=dev), torch.randn(n, 3, device=dev)

        def train():
            px.grad = None; pc.grad = None
            s, r, h = FieldFunction.apply(x, d, px, pc, st, "")
            torch.autograd.backward([s, r], [ds, dr])
        tr = prof(train)
        f = lambda t, k: f"{t.get(k, 0.0):7.1f}"
        print(f"n={n:8d} tiles={(n + 127) // 128:6d}  inference: mlp_fw {f(inf, 'field_mlp_fw_tc_kernel')} hash_fw {f(inf, 'hash_encode_fw_kernel')} | "
              f"training: mlp_fw {f(tr, 'field_mlp_fw_tc_kernel')} mlp_bw {f(tr, 'field_mlp_bw_tc_kernel')} hash_fw {f(tr, 'hash_encode_fw_kernel')} "
              f"hash_bw {f(tr, 'hash_encode_bw_runs_kernel')}  us", flush=True)


if __name__ == "__main__":
    main()
