"""Loads the UNMODIFIED reference extension built by oracle/build_ref.sh (oracle/_ref/vren*.so), if present.
Used only by the -m gpu parity tests and tests/golden/make_golden.py as ground truth."""
import glob
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_mod = None


def load():
    global _mod
    if _mod is None:
        cands = glob.glob(os.path.join(ROOT, "oracle", "_ref", "vren*.so"))
        if not cands:
            return None
        import torch  # noqa: F401  (libtorch must be loaded first)
        spec = importlib.util.spec_from_file_location("vren", cands[0])
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def canonical_train(out):
    """The reference lays rays out in atomic order (raymarching.cu:237-241).  Returns per-ray lists sorted by ray_idx:
    (n_samples (R), concatenated xyzs, dirs, deltas, ts in ray order) as CPU numpy arrays."""
    import numpy as np
    rays_a, xyzs, dirs, deltas, ts, counter = [t.cpu().numpy() for t in out]
    order = np.argsort(rays_a[:, 0], kind="stable")
    ra = rays_a[order]
    idx = np.concatenate([np.arange(s, s + n) for _, s, n in ra]) if ra[:, 2].sum() > 0 else np.zeros(0, np.int64)
    return ra[:, 2], xyzs[idx], dirs[idx], deltas[idx], ts[idx], int(counter[0])
