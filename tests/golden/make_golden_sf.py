"""Golden vectors of the shadow-field path (insert/shadow_fields.py: soft_shadow_map, SimplifySF / ComplexSF.fetch_sh) from
the UNMODIFIED reference, run on the CPU in the build container:  python tests/golden/make_golden_sf.py
Same import arrangement as make_golden_sg.py (the module-level switch to CUDA default tensors and the imports of libraries
that are absent here are neutralised; the field objects are built with __new__ because their __init__ reads git-ignored
data files and moves them to the GPU).  Writes tests/golden/shadow_field_ref.npz (outputs only; inputs come from
sg_inputs.make_sf_inputs)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from sg_inputs import make_sf_inputs  # noqa: E402

REF = "/root/reference/insert"


def main():
    for name in ("open3d", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    torch.set_default_tensor_type = lambda *_a, **_k: None  # shadow_fields.py:8
    import shadow_fields

    T = torch.from_numpy
    out = {}
    for tag, cls, k, vol_range, seed in (("simple9", shadow_fields.SimplifySF, 9, 6, 0), ("complex16", shadow_fields.ComplexSF, 16, 4, 1)):
        d = make_sf_inputs(seed, k)
        sf = object.__new__(cls)          # shadow_fields.py:82-86 / :105-108 with stand-in data, on the CPU
        sf.vol_range, sf.sh_coeff_num, sf.sf_vol = vol_range, k, T(d["sf_vol"])
        pts, pos, rot, sh = T(d["pts"]), T(d["model_pos"]), T(d["rot_inv"]), T(d["model_sh9"])
        with torch.no_grad():
            out[tag + "_sh"] = sf.fetch_sh(d["model_radius"], pts - pos).numpy()
            out[tag + "_shadow"] = shadow_fields.soft_shadow_map(sf, pos, d["model_radius"], sh, pts).numpy()          # main.py:438
            out[tag + "_shadow_rot"] = shadow_fields.soft_shadow_map(sf, pos, d["model_radius"], sh, pts, rot).numpy()  # main.py:434
    np.savez_compressed(os.path.join(HERE, "shadow_field_ref.npz"), **out)
    for k_, v in out.items():
        print(k_, v.shape, float(np.abs(v).mean()), float(v.min()), float(v.max()))


if __name__ == "__main__":
    main()
