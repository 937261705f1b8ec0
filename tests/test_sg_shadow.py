"""SG shadow / SG shading of an inserted object (insert/sg_shadow.py:103-153, insert/render_utils.py:321-375).

CPU (-m "not gpu"): the numpy oracle (oracle/sg_shadow.py) against tests/golden/sg_shadow_ref.npz -- outputs of the
UNMODIFIED reference functions (tests/golden/make_golden_sg.py) on the seeded inputs of tests/golden/sg_inputs.py.
GPU (-m gpu): libarnerf.so (arn_sg_shadow_factor / arn_sg_shade / arn_sg_shade_px through ar_nerf_b200.sg_shadow) against
the same golden vectors and, on a second seed, against the oracle.

Tolerance: 1e-4 relative (floor 1e-3 of the tensor's largest magnitude).  SG_render_core is ill-conditioned in float32
at grazing view angles -- the reference's own float32 result moves by up to 0.2 against its float64 evaluation there --
so the radiance band of a pixel is 1e-4 relative + 4 x the pixel's float32 NOISE: the largest change of the reference's
own float32 output under +-1 ulp jitter of its inputs (16 draws, stored in the golden file; the reference's float64
evaluation lies within 2 noise units of its float32 one).  The oracle-vs-CUDA comparison on the second seed measures
the noise the same way on the oracle."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from sg_inputs import make_inputs  # noqa: E402
from oracle import sg_shadow as osg  # noqa: E402

RTOL, FLOOR = 1e-4, 1e-3


def golden(tag=""):
    return np.load(os.path.join(ROOT, "tests", "golden", f"sg_shadow_ref{tag}.npz"))


# the two golden sets: (file tag, inputs, vol_range) -- the small default geometry and the insertion tool's own
# (insert/main.py:107: SGShadow(pca_path, 20, 128, 2, envH=74, envW=148))
def golden_case(tag):
    if tag == "":
        return make_inputs(0), 4
    return make_inputs(3, ncomp=128, grid=(20, 20, 20), env=(74, 148)), 2


def close(got, ref, what, band=None, rtol=RTOL):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    denom = np.maximum(np.abs(ref), FLOOR * np.abs(ref).max())
    err = np.abs(got - ref)
    if band is not None:
        err = np.maximum(err - 4.0 * band, 0.0)
    rel = err / denom
    assert rel.max() <= rtol, f"{what}: max rel err {rel.max():.3e} at {np.unravel_index(rel.argmax(), rel.shape)}"


# ------------------------------------------------------------------------------------------------------------ oracle (CPU)
def _oracle_all(d, fh, vol_range=4):
    cv = d["coeff_volume"][0]
    tabs = (cv, d["components"], d["mean"], fh)
    vr = dict(vol_range=vol_range)
    lrot = d["lSGs"].copy(); lrot[:, :3] = (d["rot_inv"] @ lrot[:, :3].T).T
    out = {}
    out["factor"] = osg.calc_shadow_factor(d["model_radius"], d["pts"], d["model_pos"], d["lSGs"], *tabs, **vr)
    out["factor_rot"] = osg.calc_shadow_factor(d["model_radius"], d["pts"], d["model_pos"], lrot, *tabs, rot_inv=d["rot_inv"], **vr)
    dec = osg.calc_self_shadow_light_decay(d["model_radius"], d["pts"], d["model_pos"], d["lSGs"], *tabs, **vr)
    dec_rot = osg.calc_self_shadow_light_decay(d["model_radius"], d["pts"], d["model_pos"], d["lSGs"], *tabs, rot_inv=d["rot_inv"], **vr)
    out["decay_full"], out["decay_rot_full"] = dec, dec_rot
    out["decay"], out["decay_rot"] = dec[:64], dec_rot[:64]
    g = [d[k] for k in ("albedo", "metal", "rough", "normal", "vdirs")]
    out["radiance_clamp"] = osg.sg_render_core(*g, dec, True, True)
    out["radiance_hdr"] = osg.sg_render_core(*g, dec_rot, False, True)
    out["radiance_noshadow"] = osg.sg_render_core(*g, d["lSGs"], False, False)
    return out


@pytest.mark.parametrize("tag", ["", "_tool"])
def test_oracle_matches_unmodified_reference(tag):
    g = golden(tag)
    d, vol_range = golden_case(tag)
    o = _oracle_all(d, golden()["fh_tab"], vol_range)
    for k in ("factor", "factor_rot", "decay", "decay_rot"):
        close(o[k], g[k], k, rtol=RTOL if tag == "" else 3e-4)   # 128-term ssdf sums next to steep f_h columns: see the second-seed test
    for k in ("radiance_clamp", "radiance_hdr", "radiance_noshadow"):
        close(o[k], g[k], k, band=g[k + "_noise"])


def test_grid_sample_rules_against_torch():
    """The oracle's grid_sample (bilinear, border) against torch.nn.functional.grid_sample, both align_corners settings."""
    import torch
    import torch.nn.functional as F
    r = np.random.RandomState(3)
    img = r.randn(3, 7, 5).astype(np.float32)
    gx, gy = r.uniform(-1.3, 1.3, 200).astype(np.float32), r.uniform(-1.3, 1.3, 200).astype(np.float32)
    for ac in (False, True):
        ref = F.grid_sample(torch.from_numpy(img)[None], torch.from_numpy(np.stack([gx, gy], -1))[None, None], padding_mode="border",
                            align_corners=ac, mode="bilinear")[0, :, 0].T.numpy()
        np.testing.assert_allclose(osg.grid_sample_2d(img, gx, gy, ac), ref, rtol=1e-5, atol=1e-6)
    vol = r.randn(4, 3, 5, 6).astype(np.float32)
    g3 = r.uniform(-1.2, 1.2, (100, 3)).astype(np.float32)
    ref = F.grid_sample(torch.from_numpy(vol)[None], torch.from_numpy(g3).reshape(1, 1, 1, -1, 3), mode="bilinear", padding_mode="border",
                        align_corners=True)[0, :, 0, 0].T.numpy()
    np.testing.assert_allclose(osg.grid_sample_3d(vol, g3, True), ref, rtol=1e-5, atol=1e-6)


def test_shadow_factor_properties():
    """A point far from the model is not shadowed more than a near one on the same ray; factors lie in [0, 1]."""
    g = golden(); d = make_inputs(0)
    f = osg.calc_shadow_factor(d["model_radius"], d["pts"], d["model_pos"], d["lSGs"], d["coeff_volume"][0], d["components"], d["mean"], g["fh_tab"])
    assert f.min() >= 0.0 and f.max() <= 1.0
    dec = osg.calc_self_shadow_light_decay(d["model_radius"], d["pts"][:8], d["model_pos"], d["lSGs"], d["coeff_volume"][0], d["components"],
                                           d["mean"], g["fh_tab"])
    assert np.array_equal(dec[..., :4], np.broadcast_to(d["lSGs"][None, :, :4], dec[..., :4].shape))  # only the colours are attenuated
    assert np.all(dec[..., 4:] <= d["lSGs"][None, :, 4:] * (1 + 1e-6))


def _jitter(x, rs):
    """Every element moved by -1 / 0 / +1 ulp (float32)."""
    x = np.ascontiguousarray(x, np.float32)
    step = rs.choice([-1, 0, 1], size=x.shape)
    return np.where(step == 0, x, np.nextafter(x, np.where(step > 0, np.float32(np.inf), np.float32(-np.inf)).astype(np.float32))).astype(np.float32)


def test_host_mirror_reads_the_reference_files_and_has_no_cpu_path(tmp_path):
    """ar_nerf_b200.sg_shadow.SGShadow.__init__ mirrors insert/sg_shadow.py:11-32: same files, same table layout, same
    attribute names; the per-pixel methods refuse host tensors (there is no CPU / PyTorch fallback)."""
    import torch
    from ar_nerf_b200.sg_shadow import SGShadow
    g, C_, eh, ew = 5, 8, 6, 7
    r = np.random.RandomState(0)
    data = {"coeff": torch.from_numpy(r.randn(g * g * g, C_).astype(np.float32)), "component": torch.from_numpy(r.randn(C_, eh, ew).astype(np.float32)),
            "mean": torch.from_numpy(r.randn(1, eh, ew).astype(np.float32))}
    pca, fh = str(tmp_path / "pca.pt"), str(tmp_path / "fh_pretab.npy")
    torch.save(data, pca); np.save(fh, r.rand(12, 9).astype(np.float32))
    sg = SGShadow(pca, g, C_, 2, envH=eh, envW=ew, fh_tab_path=fh, device="cpu")
    want = data["coeff"].reshape(g, g, g, C_).permute(3, 2, 1, 0).unsqueeze(0)       # sg_shadow.py:27-29
    assert torch.equal(sg.coeff_volume, want) and sg.coeff_volume.shape == (1, C_, g, g, g)
    assert torch.equal(sg._coeff_cl, want[0].permute(1, 2, 3, 0)) and sg._coeff_cl.is_contiguous()   # channel-last copy the kernels read
    assert torch.equal(sg.components, data["component"]) and torch.equal(sg.mean, data["mean"])
    assert sg.fh_tab.shape == (1, 1, 12, 9) and sg.vol_range == 2 and sg.ncomponents == C_ and (sg.envH, sg.envW) == (eh, ew)
    assert (sg.delta_angle_decay_fac, sg.delta_shadow_fac, sg.delta_self_shadow_fac) == (0.4, 2, 0.1)
    assert abs(float(sg.raw_h_angle) - float(np.arcsin(0.5))) < 1e-7
    lSGs = torch.from_numpy(make_inputs(0)["lSGs"])
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        sg.calc_shadow_factor(0.3, torch.zeros(4, 3), torch.zeros(3), lSGs)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        sg.calc_self_shadow_light_dacay(0.3, torch.zeros(4, 3), torch.zeros(3), lSGs)


# ------------------------------------------------------------------------------------------------------------ CUDA (GPU)
def _cuda_all(d, fh, vol_range=4):
    import torch
    from ar_nerf_b200.sg_shadow import SG_render_core, SGShadow
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sg = SGShadow.from_tensors(T(d["coeff_volume"]), T(d["components"]), T(d["mean"]), T(fh), vol_range=vol_range, device=dev)
    lSGs, pts, pos, rot = T(d["lSGs"]), T(d["pts"]), T(d["model_pos"]), T(d["rot_inv"])
    lrot = lSGs.clone(); lrot[:, :3] = (rot @ lrot[:, :3].T).T
    out = {}
    out["factor"] = sg.calc_shadow_factor(d["model_radius"], pts, pos, lSGs)
    out["factor_rot"] = sg.calc_shadow_factor(d["model_radius"], pts, pos, lrot, rot)
    dec = sg.calc_self_shadow_light_dacay(d["model_radius"], pts, pos, lSGs)
    dec_rot = sg.calc_self_shadow_light_dacay(d["model_radius"], pts, pos, lSGs, rot)
    out["decay_full"], out["decay_rot_full"] = dec, dec_rot
    out["decay"], out["decay_rot"] = dec[:64], dec_rot[:64]
    g = [T(d[k]) for k in ("albedo", "metal", "rough", "normal", "vdirs")]
    out["radiance_clamp"] = SG_render_core(*g, dec, True, True)
    out["radiance_hdr"] = SG_render_core(*g, dec_rot, False, True)
    out["radiance_noshadow"] = SG_render_core(*g, lSGs, False, False)
    # the fused form (no (px, lx, 7) tensor): main.py:559-576 in one kernel
    out["fused_clamp"] = sg.shade(d["model_radius"], pts, pos, lSGs, None, *g, True)
    out["fused_hdr"] = sg.shade(d["model_radius"], pts, pos, lSGs, rot, *g, False)
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["", "_tool"])
def test_cuda_matches_unmodified_reference(tag):
    g = golden(tag)
    d, vol_range = golden_case(tag)
    c = _cuda_all(d, golden()["fh_tab"], vol_range)
    for k in ("factor", "factor_rot", "decay", "decay_rot"):
        close(c[k], g[k], k, rtol=RTOL if tag == "" else 3e-4)
    for k in ("radiance_clamp", "radiance_hdr", "radiance_noshadow"):
        close(c[k], g[k], k, band=g[k + "_noise"])
    close(c["fused_clamp"], g["radiance_clamp"], "fused_clamp", band=g["radiance_clamp_noise"])
    close(c["fused_hdr"], g["radiance_hdr"], "fused_hdr", band=g["radiance_hdr_noise"])


@pytest.mark.gpu
def test_cuda_matches_oracle_second_seed():
    g = golden()
    # the insertion tool's geometry (insert/main.py:107: 128 components, a 74 x 148 environment map) on a non-cubic volume
    d = make_inputs(7, ncomp=128, grid=(6, 9, 7), env=(74, 148))
    o, c = _oracle_all(d, g["fh_tab"]), _cuda_all(d, g["fh_tab"])
    # The f_h table spans 200 orders of magnitude: where ssdf falls next to a steep column, the interpolation weight -- and
    # with it f_h -- carries the rounding of the 128-term ssdf sum.  Same band rule as for the radiance: 1e-4 relative + 4 x
    # the float32 noise of the oracle under +-1 ulp jitter of the inputs.
    rs0 = np.random.RandomState(11)
    jit0 = lambda x: _jitter(x, rs0)
    noise = {k: 0.0 for k in ("factor", "factor_rot", "decay_full", "decay_rot_full")}
    for _ in range(4):
        dj = dict(d)
        for key in ("pts", "lSGs", "coeff_volume", "components", "mean"):
            dj[key] = jit0(d[key])
        oj = _oracle_all(dj, g["fh_tab"])
        for k in noise:
            noise[k] = np.maximum(noise[k], np.abs(oj[k] - o[k]))
    for k in ("factor", "factor_rot", "decay_full", "decay_rot_full"):
        close(c[k], o[k], k, band=noise[k])
    # radiance: the same band rule as against the reference, with the float32 noise measured on the oracle (+-1 ulp jitter)
    rs = np.random.RandomState(5)
    jit = lambda x: _jitter(x, rs)
    names = ("albedo", "metal", "rough", "normal", "vdirs")
    for k, lights, clamp, shadow in (("radiance_clamp", o["decay_full"], True, True), ("radiance_hdr", o["decay_rot_full"], False, True),
                                     ("radiance_noshadow", d["lSGs"], False, False)):
        base = osg.sg_render_core(*[d[n] for n in names], lights, False, shadow)
        noise = np.zeros_like(base)
        for _ in range(8):
            noise = np.maximum(noise, np.abs(osg.sg_render_core(*[jit(d[n]) for n in names], jit(lights), False, shadow) - base))
        close(c[k], o[k], k, band=noise)
        if k == "radiance_clamp":
            close(c["fused_clamp"], o[k], "fused", band=noise)


@pytest.mark.gpu
def test_cuda_sg_edge_cases():
    import torch
    from ar_nerf_b200.sg_shadow import SGShadow
    g = golden(); d = make_inputs(0)
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sg = SGShadow.from_tensors(T(d["coeff_volume"]), T(d["components"]), T(d["mean"]), T(g["fh_tab"]), device=dev)
    empty = sg.calc_shadow_factor(0.3, torch.empty(0, 3, device=dev), T(d["model_pos"]), T(d["lSGs"]))
    assert empty.shape == (0,)
    one = sg.calc_shadow_factor(0.3, T(d["pts"][:1]), T(d["model_pos"]), T(d["lSGs"][:1]))      # one point, one light
    ref = osg.calc_shadow_factor(0.3, d["pts"][:1], d["model_pos"], d["lSGs"][:1], d["coeff_volume"][0], d["components"], d["mean"], g["fh_tab"])
    close(one.cpu().numpy(), ref, "one point / one light")
    with pytest.raises(RuntimeError):
        sg.calc_shadow_factor(0.3, T(d["pts"]).cpu(), T(d["model_pos"]), T(d["lSGs"]))           # host tensor: rejected like the kernels' inputs
    with pytest.raises(RuntimeError):
        sg.calc_shadow_factor(0.3, T(d["pts"]), T(d["model_pos"]), T(np.tile(d["lSGs"], (3, 1))))  # 96 lights > ARN_SG_MAX_LIGHTS
