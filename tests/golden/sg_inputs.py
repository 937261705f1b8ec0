"""Seeded inputs of the SG-shadow golden vectors (shared by make_golden_sg.py, which feeds them to the UNMODIFIED reference,
and by the tests, which feed them to the oracle and to libarnerf.so).  Everything except the f_h table -- which
make_golden_sg.py computes with the reference's own pretabulate_fh.py integrand and stores in the .npz -- is regenerated
from the seed, so that only outputs have to be committed."""
import numpy as np

GRID = 8          # PCA coefficient volume (reference: 20^3)
NCOMP = 32        # PCA components
ENV = 16          # environment-map resolution of the components (reference: 128 x 128)
N_LIGHTS = 32
N_PTS = 600
FH_LBD, FH_THETA = 64, 32   # f_h table (reference: 2048 x 1024)


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def make_inputs(seed=0, ncomp=NCOMP, grid=(GRID, GRID, GRID), env=(ENV, ENV)):
    """grid = (D, H, W) of the coefficient volume, env = (envH, envW); the golden file uses the defaults."""
    r = np.random.RandomState(seed)
    f = np.float32
    d = {}
    d["coeff_volume"] = (r.randn(1, ncomp, *grid) * 0.3 * np.sqrt(NCOMP / ncomp)).astype(f)   # 1,C,D,H,W as SGShadow.__init__ lays it out
    d["components"] = (r.randn(ncomp, *env) * 0.2).astype(f)
    d["mean"] = (r.randn(1, *env) * 0.3).astype(f)
    axis = _unit(r.randn(N_LIGHTS, 3))
    lam = 10.0 ** r.uniform(-0.5, 3.0, (N_LIGHTS, 1))
    col = r.uniform(0.05, 2.0, (N_LIGHTS, 3))
    d["lSGs"] = np.concatenate([axis, lam, col], 1).astype(f)
    # points around a model of radius 0.3 at model_pos: inside the tabulated volume, on its border and far outside
    d["model_pos"] = np.array([0.1, -0.05, 0.2], f)
    d["model_radius"] = 0.3
    rad = np.concatenate([r.uniform(0.2, 1.2, N_PTS // 2), r.uniform(1.2, 8.0, N_PTS - N_PTS // 2)])[:, None]
    d["pts"] = (d["model_pos"] + _unit(r.randn(N_PTS, 3)) * rad).astype(f)
    # a proper rotation (model_rot_inv)
    q, _ = np.linalg.qr(r.randn(3, 3))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    d["rot_inv"] = q.astype(f)
    # G-buffer of the inserted object
    d["normal"] = (_unit(r.randn(N_PTS, 3)) * r.uniform(0.5, 2.0, (N_PTS, 1))).astype(f)   # not normalised, as the caller passes it
    view = _unit(r.randn(N_PTS, 3))
    flip = np.sum(view * d["normal"], -1, keepdims=True) > 0          # most pixels face the camera (vdirs points INTO the surface)
    d["vdirs"] = np.where(flip & (r.rand(N_PTS, 1) < 0.9), -view, view).astype(f)
    d["albedo"] = r.uniform(0.0, 1.0, (N_PTS, 3)).astype(f)
    d["metal"] = r.uniform(0.0, 1.0, (N_PTS, 1)).astype(f)
    d["rough"] = r.uniform(0.2, 1.0, (N_PTS, 1)).astype(f)
    return d


def make_sf_inputs(seed=0, k=9, grid=(12, 10, 11)):
    """Shadow-field inputs (insert/shadow_fields.py): sf_vol (1,K,D,H,W) SH coefficients of the object's visibility, the
    lighting's SH model_sh9 (1,K,3), scene points around the model, a rotation."""
    r = np.random.RandomState(100 + seed)
    f = np.float32
    d = {}
    vol = r.randn(1, k, *grid) * 0.2
    vol[:, 0] = np.abs(vol[:, 0]) * 2 + 3.0                   # visibility: a dominant positive DC term (shadow ratios around 0.85 .. 1.1)
    d["sf_vol"] = vol.astype(f)
    sh = r.randn(1, k, 3) * 0.3
    sh[:, 0, :] = np.abs(sh[:, 0, :]) + 1.0
    d["model_sh9"] = sh.astype(f)
    d["model_pos"] = np.array([-0.2, 0.1, 0.05], f)
    d["model_radius"] = 0.25
    rad = np.concatenate([r.uniform(0.1, 1.0, N_PTS // 2), r.uniform(1.0, 4.0, N_PTS - N_PTS // 2)])[:, None]
    d["pts"] = (d["model_pos"] + _unit(r.randn(N_PTS, 3)) * rad).astype(f)
    q, _ = np.linalg.qr(r.randn(3, 3))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    d["rot_inv"] = q.astype(f)
    return d
