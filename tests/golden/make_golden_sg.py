"""Golden vectors of the SG-shadow path from the UNMODIFIED reference.

Run in the build container (needs /root/reference; CPU only):  python tests/golden/make_golden_sg.py
Imports insert/sg_shadow.py and insert/render_utils.py as they are -- only the module-level switch to CUDA default tensors
and the imports of libraries that are not installed here (open3d, matplotlib: unused by these functions) are neutralised
-- and calls SGShadow.calc_shadow_factor, SGShadow.calc_self_shadow_light_dacay and SG_render_core on the seeded inputs
of sg_inputs.py.  SGShadow.__init__ reads git-ignored data files and moves them to the GPU, so the object is built with
__new__ and given the same attributes from synthetic stand-ins; the f_h table is computed with the
integrand of insert/pretabulate_fh.py on a reduced grid.  Writes tests/golden/sg_shadow_ref.npz."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from sg_inputs import FH_LBD, FH_THETA, make_inputs  # noqa: E402

REF = "/root/reference/insert"


def fh_table():
    """insert/pretabulate_fh.py:5-10,28-38 (inte + the loops of pretabulate) on a FH_LBD x FH_THETA grid."""
    from scipy import integrate

    def inte(lbd, theta_d):
        def inte_func(zeta, delta):
            return np.exp(lbd * (np.sin(zeta) * np.sin(delta) - 1)) * np.sin(zeta)
        return integrate.dblquad(inte_func, np.pi / 2 - theta_d, np.pi, 0, np.pi)[0]
    theta_ds = np.linspace(-np.pi / 2, np.pi / 2, FH_THETA)
    lbds = 10 ** np.linspace(-1, 4, FH_LBD)
    res = np.ones((FH_LBD, FH_THETA), np.float32)
    for i, lbd in enumerate(lbds):
        for j, th in enumerate(theta_ds):
            res[i, j] = inte(lbd, th)
    return res


def main():
    for name in ("open3d", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    torch.set_default_tensor_type = lambda *_a, **_k: None  # sg_shadow.py:8 switches the default to CUDA tensors
    import render_utils
    import sg_shadow

    prev = os.path.join(HERE, "sg_shadow_ref.npz")
    fh = np.load(prev)["fh_tab"] if os.path.exists(prev) and "--recompute-fh" not in sys.argv else fh_table()  # 90 s of scipy dblquad
    # two geometries: the small default one, and the insertion tool's own (insert/main.py:107: SGShadow(pca, 20, 128, 2, envH=74, envW=148))
    for tag, d, vol_range in (("", make_inputs(0), 4), ("_tool", make_inputs(3, ncomp=128, grid=(20, 20, 20), env=(74, 148)), 2)):
        run_case(sg_shadow, render_utils, d, fh, vol_range, os.path.join(HERE, f"sg_shadow_ref{tag}.npz"), keep_fh=(tag == ""))


def run_case(sg_shadow, render_utils, d, fh, vol_range, path, keep_fh):
    T = torch.from_numpy
    sg = object.__new__(sg_shadow.SGShadow)  # __init__ (sg_shadow.py:11-32) with stand-in data, on the CPU
    sg.delta_angle_decay_fac, sg.delta_shadow_fac, sg.delta_self_shadow_fac = 0.4, 2, 0.1
    sg.vol_range = vol_range
    sg.raw_h_angle = torch.asin(torch.Tensor([1.0 / vol_range]))
    sg.ncomponents, sg.envH, sg.envW = d["components"].shape[0], d["components"].shape[1], d["components"].shape[2]
    sg.fh_tab = T(fh)[None, None, ...]
    sg.coeff_volume, sg.components, sg.mean = T(d["coeff_volume"]), T(d["components"]), T(d["mean"])

    lSGs, pts, pos, rot = T(d["lSGs"]), T(d["pts"]), T(d["model_pos"]), T(d["rot_inv"])
    out = {"fh_tab": fh} if keep_fh else {"vol_range": np.int32(vol_range)}
    with torch.no_grad():
        out["factor"] = sg.calc_shadow_factor(d["model_radius"], pts, pos, lSGs).numpy()
        lrot = lSGs.clone(); lrot[:, :3] = (rot @ lrot[:, :3].T).T          # main.py:496-499
        out["factor_rot"] = sg.calc_shadow_factor(d["model_radius"], pts, pos, lrot, rot).numpy()
        dec = sg.calc_self_shadow_light_dacay(d["model_radius"], pts, pos, lSGs)
        dec_rot = sg.calc_self_shadow_light_dacay(d["model_radius"], pts, pos, lSGs, rot)
        out["decay"] = dec[:64].numpy(); out["decay_rot"] = dec_rot[:64].numpy()
        args = [T(d[k]) for k in ("albedo", "metal", "rough", "normal", "vdirs")]
        out["radiance_clamp"] = render_utils.SG_render_core(*args, dec, True, True).numpy()          # main.py:572-576
        out["radiance_hdr"] = render_utils.SG_render_core(*args, dec_rot, False, True).numpy()
        out["radiance_noshadow"] = render_utils.SG_render_core(*args, lSGs, False, False).numpy()  # sg_use_self_shadow = False
        # Each pixel's float32 NOISE: the largest change of the reference's own (float32, unclamped) result when every input is
        # moved by -1 / 0 / +1 ulp, over 16 random draws.  SG_render_core is ill-conditioned in float32 at grazing view angles
        # (sharpness of the distribution SG ~ 1 / (n.v); exp(lambda * (|um| - 1)) amplifies the rounding of |um|): there the
        # reference's output is rounding noise (its float64 evaluation lies up to 0.2 away, i.e. ~2 noise units).  The tests
        # hold a result to |x - reference| <= 1e-4 relative + 4 noise units.
        rs = np.random.RandomState(1)
        def jit(x):
            x = x.numpy(); step = rs.choice([-1, 0, 1], size=x.shape)
            far = np.where(step > 0, np.float32(np.inf), np.float32(-np.inf)).astype(np.float32)
            return torch.from_numpy(np.where(step == 0, x, np.nextafter(x, far)).astype(np.float32))
        for name, lights, shadow in (("radiance_clamp", dec, True), ("radiance_hdr", dec_rot, True), ("radiance_noshadow", lSGs, False)):
            base = render_utils.SG_render_core(*args, lights, False, shadow).numpy()
            noise = np.zeros_like(base)
            for _ in range(16):
                noise = np.maximum(noise, np.abs(render_utils.SG_render_core(*[jit(a) for a in args], jit(lights.contiguous()), False, shadow).numpy() - base))
            out[name + "_noise"] = noise
    np.savez_compressed(path, **out)
    print(path)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).mean()))


if __name__ == "__main__":
    main()
