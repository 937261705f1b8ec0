"""SG shadow and SG shading of an inserted object on libarnerf.so -- host-side mirror of insert/sg_shadow.py (class SGShadow)
and of insert/render_utils.py:321-375 SG_render_core, same names, arguments and results.

    sg = SGShadow(pca_path, ...)                       # as the reference: PCA file + ./insert/data/fh_pretab.npy
    sg = SGShadow.from_tensors(coeff_volume, components, mean, fh_tab, ...)
    smap = sg.calc_shadow_factor(model_r, pts, model_pos, lSGs, rot_inv)            # (px,)
    lSGs_px = sg.calc_self_shadow_light_dacay(model_r, pts, model_pos, lSGs, rot)   # (px, lx, 7)
    cols = SG_render_core(albedo, metal, rough, normal, vdirs, lSGs_px, clamp01, self_shadow=True)
    cols = sg.shade(model_r, pts, model_pos, lSGs, rot, albedo, metal, rough, normal, vdirs, clamp01)   # the two fused: no (px,lx,7)

There is no PyTorch fallback: every per-pixel result comes from arn_sg_shadow_factor / arn_sg_shade."""
import ctypes as C

import numpy as np
import torch

from ._lib import SgTables, call, check_tensor, ptr, stream


def _f32(t, device):
    return torch.as_tensor(t, dtype=torch.float32, device=device).contiguous()


class SGShadow:
    def __init__(self, pca_path, grid_size=20, ncomponents=32, vol_range=4, envH=128, envW=128,
                 angle_decay_fac=0.4, shadow_pow_fac=2, self_shadow_pow_fac=0.1, fh_tab_path='./insert/data/fh_pretab.npy', device='cuda'):
        """sg_shadow.py:11-32 (the tables are read from the same files and laid out the same way)."""
        data = torch.load(pca_path, map_location='cpu')
        coeff_volume = data['coeff'].reshape(grid_size, grid_size, grid_size, ncomponents).permute(3, 2, 1, 0).unsqueeze(0)  # 1,C,D,H,W
        self._setup(coeff_volume, data['component'], data['mean'], torch.from_numpy(np.load(fh_tab_path)), vol_range, envH, envW,
                    angle_decay_fac, shadow_pow_fac, self_shadow_pow_fac, device)

    @classmethod
    def from_tensors(cls, coeff_volume, components, mean, fh_tab, vol_range=4, angle_decay_fac=0.4, shadow_pow_fac=2,
                     self_shadow_pow_fac=0.1, device='cuda'):
        """coeff_volume (1,C,D,H,W), components (C,envH,envW), mean (1,envH,envW), fh_tab (sharpness rows, angle columns)."""
        self = object.__new__(cls)
        self._setup(coeff_volume, components, mean, fh_tab, vol_range, components.shape[1], components.shape[2], angle_decay_fac,
                    shadow_pow_fac, self_shadow_pow_fac, device)
        return self

    def _setup(self, coeff_volume, components, mean, fh_tab, vol_range, envH, envW, angle_decay_fac, shadow_pow_fac, self_shadow_pow_fac, device):
        dev = torch.device(device)
        self.delta_angle_decay_fac, self.delta_shadow_fac, self.delta_self_shadow_fac = angle_decay_fac, shadow_pow_fac, self_shadow_pow_fac
        self.vol_range = vol_range
        self.raw_h_angle = torch.asin(torch.tensor([1.0 / vol_range]))
        self.coeff_volume = _f32(coeff_volume, dev)                        # 1,C,D,H,W (reference layout, kept as an attribute)
        self.components = _f32(components, dev)                            # C,envH,envW
        self.mean = _f32(mean, dev).reshape(1, envH, envW)
        self.fh_tab = _f32(fh_tab, dev).reshape(1, 1, *fh_tab.shape[-2:])  # 1,1,rows,cols
        self.ncomponents, self.envH, self.envW = self.components.shape[0], envH, envW
        _, Cn, D, H, W = self.coeff_volume.shape
        self._coeff_cl = self.coeff_volume[0].permute(1, 2, 3, 0).contiguous()  # D,H,W,C: one corner = C contiguous floats
        self._tables = SgTables(ptr(self._coeff_cl), D, H, W, Cn, ptr(self.components), ptr(self.mean), envH, envW,
                                ptr(self.fh_tab), self.fh_tab.shape[2], self.fh_tab.shape[3], float(vol_range), float(angle_decay_fac),
                                float(shadow_pow_fac), float(self_shadow_pow_fac))
        self._scratch = None

    # ---------------------------------------------------------------------------------------------------------------
    def _light_scratch(self, n_lights, device):
        need = n_lights * (self.ncomponents + 12) + 3
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != device:
            self._scratch = torch.empty(need, dtype=torch.float32, device=device)
        return self._scratch

    @staticmethod
    def _host3(v):
        return (C.c_float * 3)(*[float(x) for x in torch.as_tensor(v).reshape(-1).tolist()])

    @staticmethod
    def _host9(m):
        return None if m is None else (C.c_float * 9)(*[float(x) for x in torch.as_tensor(m).reshape(-1).tolist()])

    def _check_lights(self, lSGs):
        lSGs = check_tensor(lSGs.contiguous().float(), "lSGs", torch.float32, 2, 7)
        if lSGs.shape[0] > 64:
            raise RuntimeError("at most 64 SG lights")
        return lSGs

    def calc_shadow_factor(self, scale, pts, model_pos, lSGs, rot_inv=None):
        """sg_shadow.py:103-116.  pts (px,3); lSGs (lx,7), already rotated by the caller when rot_inv is given."""
        pts = check_tensor(pts.contiguous().float(), "pts", torch.float32, 2, 3)
        lSGs = self._check_lights(lSGs)
        out = torch.empty(pts.shape[0], dtype=torch.float32, device=pts.device)
        pos, rot = self._host3(model_pos), self._host9(rot_inv)
        call("arn_sg_shadow_factor", C.byref(self._tables), ptr(lSGs), lSGs.shape[0], ptr(pts), pts.shape[0], pos, rot, float(scale),
             ptr(self._light_scratch(lSGs.shape[0], pts.device)), ptr(out), stream())
        return out

    def _rotated_axes(self, lSGs, rot_inv):
        if rot_inv is None:
            return lSGs
        l_rot = lSGs.clone()                                               # sg_shadow.py:124-126
        l_rot[:, :3] = (torch.as_tensor(rot_inv, dtype=torch.float32, device=lSGs.device) @ l_rot[:, :3].T).T
        return l_rot.contiguous()

    def calc_self_shadow_light_dacay(self, scale, pts, model_pos, lSGs, rot_inv=None):
        """sg_shadow.py:118-153 -> (px, lx, 7).  (shade() fuses this with SG_render_core and never materialises it.)"""
        pts = check_tensor(pts.contiguous().float(), "pts", torch.float32, 2, 3)
        lSGs = self._check_lights(lSGs)
        out = torch.empty(pts.shape[0], lSGs.shape[0], 7, dtype=torch.float32, device=pts.device)
        call("arn_sg_shade", C.byref(self._tables), ptr(lSGs), ptr(self._rotated_axes(lSGs, rot_inv)), lSGs.shape[0], ptr(pts), pts.shape[0],
             self._host3(model_pos), self._host9(rot_inv), float(scale), None, None, None, None, None, 0, 1,
             ptr(self._light_scratch(lSGs.shape[0], pts.device)), ptr(out), None, stream())
        return out

    def shade(self, scale, pts, model_pos, lSGs, rot_inv, albedo, metal, rough, normal, vdirs, clamp01):
        """main.py:559-576 in one kernel: SG_render_core(albedo, ..., calc_self_shadow_light_dacay(...), clamp01, True)."""
        pts = check_tensor(pts.contiguous().float(), "pts", torch.float32, 2, 3)
        lSGs = self._check_lights(lSGs)
        n = pts.shape[0]
        g = [check_tensor(t.contiguous().float(), nm) for t, nm in ((albedo, "albedo"), (metal, "metal"), (rough, "rough"), (normal, "normal"), (vdirs, "vdirs"))]
        for t, k in zip(g, (3, 1, 1, 3, 3)):
            if t.numel() != n * k:
                raise RuntimeError("G-buffer tensors must have one row per point")
        out = torch.empty(n, 3, dtype=torch.float32, device=pts.device)
        call("arn_sg_shade", C.byref(self._tables), ptr(lSGs), ptr(self._rotated_axes(lSGs, rot_inv)), lSGs.shape[0], ptr(pts), n,
             self._host3(model_pos), self._host9(rot_inv), float(scale), *[ptr(t) for t in g], 1 if clamp01 else 0, 1,
             ptr(self._light_scratch(lSGs.shape[0], pts.device)), None, ptr(out), stream())
        return out


def SG_render_core(albedo, metal, rough, normal, vdirs, lSGs, clamp01, self_shadow=True, refl_probe=None, only_spec=False):
    """render_utils.py:321-375.  lSGs (px,lx,7) with self_shadow (already attenuated per pixel), (lx,7) without."""
    n = normal.shape[0]
    dev = normal.device
    g = [check_tensor(t.contiguous().float(), nm) for t, nm in ((albedo, "albedo"), (metal, "metal"), (rough, "rough"), (normal, "normal"), (vdirs, "vdirs"))]
    out = torch.empty(n, 3, dtype=torch.float32, device=dev)
    if self_shadow:  # per-pixel lights are an INPUT here (already attenuated): (px, lx, 7)
        lSGs = check_tensor(lSGs.contiguous().float(), "lSGs", torch.float32, 3, 7)
        if lSGs.shape[0] != n:
            raise RuntimeError("lSGs must hold one set of lights per pixel")
        n_lights, per_pixel = lSGs.shape[1], 1
    else:
        lSGs = check_tensor(lSGs.contiguous().float(), "lSGs", torch.float32, 2, 7)
        n_lights, per_pixel = lSGs.shape[0], 0
    call("arn_sg_shade_px", ptr(lSGs), n_lights, per_pixel, n, *[ptr(t) for t in g], 1 if clamp01 else 0, ptr(out), stream())
    return out
