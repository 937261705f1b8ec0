"""CPU oracle of the SG-shadow path -- TEST INFRASTRUCTURE, never imported by the product.

numpy (float32) restatement of
  insert/sg_shadow.py:35-53   SGShadow.light_axis_to_cood
  insert/sg_shadow.py:55-68   SGShadow.calc_inte_L_V           :70-74 calc_inte_L
  insert/sg_shadow.py:80-101  SGShadow.fetch_ssdf
  insert/sg_shadow.py:103-116 SGShadow.calc_shadow_factor      :118-153 calc_self_shadow_light_dacay
  insert/render_utils.py:266-278 SGProduct  :280-300 SGHemisphereIntegral  :304-318 SGIrradiance  :321-375 SG_render_core
  (+ the helpers at :9-13,51-61,68-72,191-192) and of torch.nn.functional.grid_sample (bilinear, padding_mode='border').
Pinned: tests/test_oracle_golden.py holds it to tests/golden/sg_shadow_ref.npz, the outputs of the unmodified reference
functions run on the CPU by tests/golden/make_golden_sg.py."""
import numpy as np

F = np.float32
PI = F(np.pi)
EPS = F(1e-6)


def _unnorm(x, size, align_corners):
    x = x.astype(F)
    if align_corners:
        return (x + F(1)) / F(2) * F(size - 1)
    return ((x + F(1)) * F(size) - F(1)) / F(2)


def grid_sample_2d(img, gx, gy, align_corners=False):
    """img (C,H,W); gx, gy (...): bilinear, border padding.  Returns (..., C)."""
    C, H, W = img.shape
    ix = np.clip(_unnorm(gx, W, align_corners), F(0), F(W - 1)); iy = np.clip(_unnorm(gy, H, align_corners), F(0), F(H - 1))
    x0 = np.floor(ix); y0 = np.floor(iy)
    wx1 = ix - x0; wy1 = iy - y0; wx0 = F(1) - wx1; wy0 = F(1) - wy1
    x0 = x0.astype(np.int64); y0 = y0.astype(np.int64)
    out = np.zeros(gx.shape + (C,), F)
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xx, yy = x0 + dx, y0 + dy
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            v = img[:, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]          # C, ...
            out += np.moveaxis(v, 0, -1) * (np.where(ok, wx * wy, F(0)))[..., None]
    return out


def grid_sample_3d(vol, g, align_corners=True):
    """vol (C,D,H,W); g (n,3) with x -> W, y -> H, z -> D: trilinear, border padding.  Returns (n, C)."""
    C, D, H, W = vol.shape
    idx = [np.clip(_unnorm(g[:, k], s, align_corners), F(0), F(s - 1)) for k, s in ((0, W), (1, H), (2, D))]
    base = [np.floor(i) for i in idx]
    w1 = [i - b for i, b in zip(idx, base)]
    w0 = [F(1) - w for w in w1]
    base = [b.astype(np.int64) for b in base]
    out = np.zeros((g.shape[0], C), F)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xx, yy, zz = base[0] + dx, base[1] + dy, base[2] + dz
                ok = (xx < W) & (yy < H) & (zz < D)
                w = (w1[0] if dx else w0[0]) * (w1[1] if dy else w0[1]) * (w1[2] if dz else w0[2])
                v = vol[:, np.clip(zz, 0, D - 1), np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]   # C, n
                out += v.T * np.where(ok, w, F(0))[:, None]
    return out


def light_axis_tables(lSGs, components, mean):
    """sg_shadow.py:35-53 -> components_s (lx, C), mean_s (lx)."""
    phi = np.arccos(lSGs[:, 1].astype(F)); theta = np.arctan2(lSGs[:, 2].astype(F), lSGs[:, 0].astype(F))
    phi_n = phi / PI * F(2) - F(1); theta_n = theta / PI
    comp = grid_sample_2d(components, theta_n, phi_n)          # x = theta (W), y = phi (H)
    mean_s = grid_sample_2d(mean, theta_n, phi_n)[:, 0]
    return comp.astype(F), mean_s.astype(F)


def fetch_ssdf(scale, m2pts, coeff_volume, comp_s, mean_s, vol_range=4, angle_decay_fac=0.4):
    """sg_shadow.py:80-101; coeff_volume (C,D,H,W)."""
    p = m2pts.astype(F) / F(scale) / F(vol_range)
    dis = np.maximum(np.linalg.norm(p, axis=-1, keepdims=True).astype(F), F(1))
    p = p / dis
    raw = np.arcsin(F(1.0 / vol_range)).astype(F)
    delta = (raw - np.arcsin(F(1) / (dis * F(vol_range)))) * F(angle_decay_fac)       # px,1
    pca = grid_sample_3d(coeff_volume, p, align_corners=True)                        # px,C
    return (pca @ comp_s.T + mean_s[None, :] + delta).astype(F)


def _fh_lookup(ssdf_clipped, lSGs, fh_tab):
    """sg_shadow.py:55-64 / :131-140: f_h table fetch at (ssdf / (pi/2), log-lambda)."""
    s = ssdf_clipped / (PI / F(2))
    lam = (np.log10(np.abs(lSGs[:, 3].astype(F) + F(1e-6))) - F(1.5)) / F(2.5)
    lam = np.broadcast_to(lam[None, :], s.shape)
    return grid_sample_2d(fh_tab[None], s, lam)[..., 0]                              # px,lx


def _m2pts(pts, model_pos, rot_inv):
    m = pts.astype(F) - model_pos.astype(F)[None, :]
    if rot_inv is not None:
        m = (rot_inv.astype(F) @ m.T).T
    return m


def calc_shadow_factor(scale, pts, model_pos, lSGs, coeff_volume, components, mean, fh_tab, rot_inv=None, vol_range=4,
                       angle_decay_fac=0.4, shadow_pow_fac=2):
    """sg_shadow.py:103-116 (lSGs already rotated by the caller when rot_inv is given, main.py:496-499)."""
    comp_s, mean_s = light_axis_tables(lSGs, components, mean)
    ssdf = np.clip(fetch_ssdf(scale, _m2pts(pts, model_pos, rot_inv), coeff_volume, comp_s, mean_s, vol_range, angle_decay_fac),
                   -PI / F(2), PI / F(2))
    cols = _fh_lookup(ssdf, lSGs, fh_tab) @ lSGs[:, 4:7].astype(F)                    # px,3
    lam = lSGs[:, 3:4].astype(F)
    inte_L = np.sum(F(2) * PI * (lSGs[:, 4:7].astype(F) / lam) * (F(1) - np.exp(-lam)), 0, keepdims=True)
    fac = np.clip(np.abs(cols / inte_L), 0, 1)
    fac = F(0.2989) * fac[:, 0] + F(0.5870) * fac[:, 1] + F(0.1140) * fac[:, 2]
    return np.power(fac, F(shadow_pow_fac)).astype(F)


def calc_self_shadow_light_decay(scale, pts, model_pos, lSGs, coeff_volume, components, mean, fh_tab, rot_inv=None, vol_range=4,
                                 angle_decay_fac=0.4, self_shadow_pow_fac=0.1):
    """sg_shadow.py:118-153 -> (px, lx, 7): the lights with their colours attenuated per pixel."""
    l_axis = lSGs.astype(F).copy()
    if rot_inv is not None:
        l_axis[:, :3] = (rot_inv.astype(F) @ l_axis[:, :3].T).T
    comp_s, mean_s = light_axis_tables(l_axis, components, mean)
    ssdf = np.clip(fetch_ssdf(scale, _m2pts(pts, model_pos, rot_inv), coeff_volume, comp_s, mean_s, vol_range, angle_decay_fac),
                   -PI / F(2), PI / F(2))
    fhs = _fh_lookup(ssdf, lSGs, fh_tab)
    lam = lSGs[:, 3].astype(F)
    fh_ns = F(2) * PI / lam * (F(1) - np.exp(-lam))
    decay = np.power(np.clip(np.abs(fhs / fh_ns[None, :]), 0, 1), F(self_shadow_pow_fac))
    out = np.broadcast_to(lSGs.astype(F)[None], (pts.shape[0],) + lSGs.shape).copy()
    out[..., 4:7] *= decay[..., None]
    return out


def sg_product(a, b):
    """render_utils.py:266-278."""
    lm = a[..., 3:4] + b[..., 3:4]
    um = (a[..., 3:4] * a[..., :3] + b[..., 3:4] * b[..., :3]) / lm
    ul = np.linalg.norm(um, axis=-1, keepdims=True).astype(F)
    out = np.ones(np.broadcast(a, b).shape, F)
    out[..., :3] = um * (F(1) / ul)
    out[..., 3:4] = lm * ul
    out[..., 4:7] = a[..., 4:7] * b[..., 4:7] * np.exp(lm * (ul - F(1)))
    return out


def sg_hemisphere_integral(sgs, normal):
    """render_utils.py:280-300."""
    cos_b = np.sum(sgs[..., :3] * normal, -1, keepdims=True)
    lam = np.maximum(sgs[..., 3:4], EPS)
    il = F(1) / lam
    t = np.sqrt(lam) * (F(1.6988) + F(10.8438) * il) / (F(1) + F(6.2201) * il + F(10.2415) * il * il)
    inv_a = np.exp(-t)
    mask = (cos_b >= 0).astype(F)
    inv_b = np.exp(-t * np.maximum(cos_b, F(0)))
    s1 = (F(1) - inv_a * inv_b) / (F(1) - inv_a + inv_b - inv_a * inv_b)
    b = np.exp(t * np.minimum(cos_b, F(0)))
    s2 = (b - inv_a) / ((F(1) - inv_a) * (b + F(1)))
    s = mask * s1 + (F(1) - mask) * s2
    A_b = F(2) * PI / lam * (np.exp(-lam) - np.exp(F(-2) * lam))
    A_u = F(2) * PI / lam * (F(1) - np.exp(-lam))
    return (A_b * (F(1) - s) + A_u * s) * sgs[..., 4:7]


def sg_irradiance(sgs, normal):
    """render_utils.py:304-318 (sum over the lights, then relu)."""
    px, lx = sgs.shape[0], sgs.shape[1]
    cos_sg = np.ones((px, 7), F); cos_sg[:, :3] = normal; cos_sg[:, 3:4] *= F(0.0315); cos_sg[:, 4:7] *= F(32.7080)
    cos_sg = np.broadcast_to(cos_sg[:, None, :], (px, lx, 7))
    n = np.broadcast_to(normal[:, None, :], (px, lx, 3))
    with np.errstate(all="ignore"):
        irr = sg_hemisphere_integral(sg_product(sgs, cos_sg), n) - F(31.7003) * sg_hemisphere_integral(sgs, n)
    return np.maximum(np.sum(irr, 1), F(0))


def sg_render_core(albedo, metal, rough, normal, vdirs, lSGs, clamp01, self_shadow=True):
    """render_utils.py:321-375.  lSGs (px,lx,7) with self_shadow, (lx,7) without."""
    albedo, metal, rough = albedo.astype(F), metal.astype(F), rough.astype(F)
    v = -vdirs.astype(F)
    n = normal.astype(F) / np.linalg.norm(normal.astype(F), axis=-1, keepdims=True).astype(F)
    px = n.shape[0]
    ndv = np.sum(n * v, -1, keepdims=True)
    D = np.ones((px, 7), F)
    D[:, :3] = ndv * n * F(2) - v                                   # reflect_dir :191-192
    m2 = rough ** 2
    D[:, 3:4] = F(2) / m2 / (F(4) * np.maximum(ndv, EPS))            # pos_dot_eps :12-13
    D[:, 4:7] *= F(1) / (PI * m2)
    L = lSGs.astype(F) if self_shadow else np.broadcast_to(lSGs.astype(F)[None], (px,) + lSGs.shape)
    Dx = np.broadcast_to(D[:, None, :], L.shape)
    with np.errstate(all="ignore"):
        spec_irr = sg_irradiance(sg_product(Dx, L), n)
        diff_irr = sg_irradiance(L, n)
        NdotV = np.maximum(ndv, F(0)); NdotL = NdotV
        F0 = np.ones_like(albedo) * F(0.04) * (F(1) - metal) + albedo * metal
        Fr = F0 + (F(1) - F0) * np.power(F(1) - NdotV, F(5))
        a = rough ** 2
        sq = a * np.maximum(F(1) / NdotV ** 2 - F(1), F(0))
        G = F(1) / (F(0.5) * (np.sqrt(F(1) + sq) - F(1)) * F(2) + F(1))
        Moi = Fr * G / (F(4) * NdotL * NdotV + EPS)
        spec = Moi * spec_irr
        diff = albedo / PI * diff_irr
        kS = F0 + (np.maximum(np.broadcast_to(F(1) - rough, F0.shape), F0) - F0) * np.power(F(1) - NdotV, F(5))
        kD = (F(1) - kS) * (F(1) - metal)
        rad = kD * diff + spec
    return (np.clip(rad, 0, 1) if clamp01 else np.maximum(rad, F(0))).astype(F)


# ---------------------------------------------------------------------------------------------------------------------
# Shadow field (the SH alternative to the SG shadow): insert/shadow_fields.py:59-78 soft_shadow_map, :92-101 / :112-121
# fetch_sh, insert/insert_utils.py:153-154 SH_product0.  Pinned by tests/golden/shadow_field_ref.npz
# (tests/golden/make_golden_sf.py: outputs of the unmodified reference functions).
def sf_fetch_sh(scale, m2pts, sf_vol, vol_range):
    """shadow_fields.py:92-101: sf_vol (K,D,H,W); points relative to the model, NOT normalised beyond the volume (border)."""
    p = m2pts.astype(F) / F(scale) / F(vol_range)
    return grid_sample_3d(sf_vol, p, align_corners=True)


def soft_shadow_map(sf_vol, vol_range, model_pos, model_r, model_sh9, pts, rot_inv=None):
    """shadow_fields.py:59-78.  model_sh9 (1,K,3) -> (px,)."""
    sh = sf_fetch_sh(model_r, _m2pts(pts, model_pos, rot_inv), sf_vol, vol_range)               # px,K
    new_ir = F(0.2821) * (sh @ model_sh9[0].astype(F))                                           # SH_product0 per colour: px,3
    old_ir = model_sh9[:, 0, :].astype(F)                                                        # 1,3
    res = np.mean(np.clip(new_ir / old_ir, 0.0, 1.0), axis=-1, dtype=F)
    return np.power(res, F(10)).astype(F)
