"""Shadow field (the SH alternative to the SG shadow) on libarnerf.so -- host-side mirror of insert/shadow_fields.py:
SimplifySF / ComplexSF (same constructor arguments, `fetch_sh`) and `soft_shadow_map` (same signature and result).
No PyTorch fallback: every per-pixel value comes from arn_sf_soft_shadow."""
import ctypes as C

import torch

from ._lib import call, check_tensor, ptr, stream


class _ShadowField:
    vol_range = 4

    def _setup(self, sf_vol, sh_coeff_num, device):
        self.sh_coeff_num = sh_coeff_num
        self.sf_vol = torch.as_tensor(sf_vol, dtype=torch.float32, device=torch.device(device)).contiguous()  # 1,K,D,H,W (reference layout)
        if self.sf_vol.dim() != 5 or self.sf_vol.shape[0] != 1 or self.sf_vol.shape[1] != sh_coeff_num or sh_coeff_num > 16:
            raise RuntimeError("sf_vol must be (1, sh_coeff_num <= 16, D, H, W)")
        self._sf_cl = self.sf_vol[0].permute(1, 2, 3, 0).contiguous()                                     # D,H,W,K: a corner = K contiguous floats

    def _run(self, scale, pts, model_pos, rot_inv, model_sh, want_sh):
        pts = check_tensor(pts.contiguous().float(), "pts", torch.float32, 2, 3)
        n, K = pts.shape[0], self.sh_coeff_num
        _, _, D, H, W = self.sf_vol.shape
        out = torch.empty((n, K) if want_sh else (n,), dtype=torch.float32, device=pts.device)
        pos = (C.c_float * 3)(*[float(x) for x in torch.as_tensor(model_pos).reshape(-1).tolist()])
        rot = None if rot_inv is None else (C.c_float * 9)(*[float(x) for x in torch.as_tensor(rot_inv).reshape(-1).tolist()])
        msh = None
        if model_sh is not None:
            m = torch.as_tensor(model_sh, dtype=torch.float32).reshape(-1, 3)
            if m.shape[0] != K:
                raise RuntimeError("model_sh9 must be (1, sh_coeff_num, 3)")
            msh = (C.c_float * (3 * K))(*[float(x) for x in m.reshape(-1).tolist()])
        call("arn_sf_soft_shadow", ptr(self._sf_cl), D, H, W, K, float(self.vol_range), msh, ptr(pts), n, pos, rot, float(scale),
             ptr(out) if want_sh else None, None if want_sh else ptr(out), stream())
        return out

    def fetch_sh(self, scale, pts):
        """shadow_fields.py:92-101 / :112-121: pts (x,3) relative to the model -> (x, sh_coeff_num)."""
        return self._run(scale, pts, (0.0, 0.0, 0.0), None, None, True)


class SimplifySF(_ShadowField):
    def __init__(self, sh_coeff_num=9, sf_path='./insert/data/sf.tar', device='cuda'):
        """shadow_fields.py:82-86: the sphere's field, stored XYZ x K on disk."""
        self.vol_range = 6
        self._setup(torch.load(sf_path, map_location='cpu').permute(3, 2, 1, 0).unsqueeze(0), sh_coeff_num, device)


class ComplexSF(_ShadowField):
    def __init__(self, sh_path, sh_coeff_num=9, device='cuda'):
        """shadow_fields.py:105-108: a field already stored as (1,K,D,H,W) (transform_sf_txt_to_torch, :48-51)."""
        self.vol_range = 4
        self._setup(torch.load(sh_path, map_location='cpu'), sh_coeff_num, device)


def soft_shadow_map(sfer, model_pos, model_r, model_sh9, pts, rot_inv=None):
    """shadow_fields.py:59-78: model_sh9 (1,K,3), pts (x,3) -> (x,) shadow factors."""
    return sfer._run(model_r, pts, model_pos, rot_inv, model_sh9, False)
