"""Shadow field (insert/shadow_fields.py:59-78 soft_shadow_map, :92-121 fetch_sh), the SH alternative to the SG shadow.

CPU: the numpy oracle (oracle/sg_shadow.py) against tests/golden/shadow_field_ref.npz -- outputs of the UNMODIFIED
reference functions (tests/golden/make_golden_sf.py).  GPU: arn_sf_soft_shadow through ar_nerf_b200.shadow_fields against
the same vectors.  1e-4 relative (the final pow(., 10) multiplies the relative rounding of the ratio by ten: observed 2e-6)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from sg_inputs import make_sf_inputs  # noqa: E402
from oracle import sg_shadow as osg  # noqa: E402

CASES = [("simple9", 9, 6, 0), ("complex16", 16, 4, 1)]   # (golden tag, SH coefficients, vol_range, seed): SimplifySF / ComplexSF


def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "shadow_field_ref.npz"))


def close(got, ref, what, rtol=1e-4):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max())
    assert rel.max() <= rtol, f"{what}: max rel err {rel.max():.3e}"


@pytest.mark.parametrize("tag,k,vol_range,seed", CASES)
def test_oracle_matches_unmodified_reference(tag, k, vol_range, seed):
    g, d = golden(), make_sf_inputs(seed, k)
    vol = d["sf_vol"][0]
    close(osg.sf_fetch_sh(d["model_radius"], d["pts"] - d["model_pos"], vol, vol_range), g[tag + "_sh"], "fetch_sh", rtol=1e-5)
    close(osg.soft_shadow_map(vol, vol_range, d["model_pos"], d["model_radius"], d["model_sh9"], d["pts"]), g[tag + "_shadow"], "shadow")
    close(osg.soft_shadow_map(vol, vol_range, d["model_pos"], d["model_radius"], d["model_sh9"], d["pts"], d["rot_inv"]), g[tag + "_shadow_rot"],
          "shadow_rot")
    s = g[tag + "_shadow"]
    assert s.min() >= 0.0 and s.max() <= 1.0 and (s == 1.0).any() and (s < 0.5).any()   # the vectors cover the clamp and the deep shadow


def test_host_mirror_layout_and_no_cpu_path(tmp_path):
    import torch
    from ar_nerf_b200.shadow_fields import ComplexSF, SimplifySF, soft_shadow_map
    d = make_sf_inputs(0, 9)
    xyzk = torch.from_numpy(d["sf_vol"][0]).permute(3, 2, 1, 0).contiguous()          # what gen_sf_3d saves: X,Y,Z,K (shadow_fields.py:42-44)
    p1, p2 = str(tmp_path / "sf.tar"), str(tmp_path / "obj_sh.tar")
    torch.save(xyzk, p1); torch.save(torch.from_numpy(d["sf_vol"]), p2)
    a, b = SimplifySF(9, sf_path=p1, device="cpu"), ComplexSF(p2, 9, device="cpu")
    assert a.vol_range == 6 and b.vol_range == 4 and a.sh_coeff_num == b.sh_coeff_num == 9
    assert torch.equal(a.sf_vol, torch.from_numpy(d["sf_vol"])) and torch.equal(b.sf_vol, a.sf_vol)
    assert torch.equal(a._sf_cl, a.sf_vol[0].permute(1, 2, 3, 0)) and a._sf_cl.is_contiguous()
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        soft_shadow_map(a, torch.zeros(3), 0.3, torch.from_numpy(d["model_sh9"]), torch.zeros(4, 3))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        a.fetch_sh(0.3, torch.zeros(4, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("tag,k,vol_range,seed", CASES)
def test_cuda_matches_unmodified_reference(tag, k, vol_range, seed):
    import torch
    from ar_nerf_b200.shadow_fields import _ShadowField, soft_shadow_map
    g, d = golden(), make_sf_inputs(seed, k)
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sf = _ShadowField(); sf.vol_range = vol_range; sf._setup(T(d["sf_vol"]), k, dev)
    pts, pos, rot, msh = T(d["pts"]), T(d["model_pos"]), T(d["rot_inv"]), T(d["model_sh9"])
    close(sf.fetch_sh(d["model_radius"], pts - pos).cpu().numpy(), g[tag + "_sh"], "fetch_sh", rtol=1e-5)
    close(soft_shadow_map(sf, pos, d["model_radius"], msh, pts).cpu().numpy(), g[tag + "_shadow"], "shadow")
    close(soft_shadow_map(sf, pos, d["model_radius"], msh, pts, rot).cpu().numpy(), g[tag + "_shadow_rot"], "shadow_rot")
    assert soft_shadow_map(sf, pos, d["model_radius"], msh, torch.empty(0, 3, device=dev)).shape == (0,)
