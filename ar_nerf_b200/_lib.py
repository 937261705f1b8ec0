"""ctypes loader for libarnerf.so (the C ABI declared in include/arnerf.h).

There is NO fallback: if the CUDA library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libarnerf.so")

P = C.c_void_p
I = C.c_int
L = C.c_int64
F = C.c_float


class Levels(C.Structure):
    _fields_ = [("scale_host", P), ("res_host", P), ("size_host", P), ("offset_host", P)]


FIELD_SCRATCH_BYTES = 20480 + 1280 * 10240 * 4  # ARN_FIELD_SCRATCH_BYTES (include/arnerf.h)


class FieldWs(C.Structure):
    _fields_ = [("feat", P), ("hid", P), ("h", P), ("in32", P), ("hid1", P), ("hid2", P), ("wimg", P)]


class TrainCfg(C.Structure):
    """Mirror of arn_train_t (include/arnerf.h) -- field order and types must match exactly."""
    _fields_ = [("rays_o", P), ("rays_d", P), ("rgb_target", P), ("noise", P), ("n_rays", L),
                ("density_bitfield", P), ("cascades", I), ("grid_size", I), ("scale", F), ("exp_step_factor", F), ("max_samples", I),
                ("T_threshold", F), ("near", F),
                ("center_host", P), ("half_size_host", P), ("xyz_min_host", P), ("xyz_max_host", P),
                ("levels", Levels), ("params_xyz_f16", P), ("params_rgb_f16", P), ("rgb_act", I),
                ("bg_host", P), ("lambda_opacity", F), ("lambda_depth", F), ("grad_scale", F), ("loss_scale", F),
                ("hits_t", P), ("rays_a", P), ("counter", P), ("t_scratch", P), ("count_scratch", P), ("total_samples", P),
                ("opacity", P), ("depth", P), ("rgb", P), ("rgb_final", P), ("dL_dopacity", P), ("dL_ddepth", P), ("dL_drgb", P),
                ("capacity", L), ("xyzs", P), ("dirs", P), ("deltas", P), ("ts", P), ("sigmas", P), ("rgbs", P), ("ws_out", P),
                ("dL_dsigmas", P), ("dL_drgbs", P), ("dfeat", P), ("ws", FieldWs),
                ("grad_xyz", P), ("grad_rgb", P), ("loss_out", P)]


class TestIterCfg(C.Structure):
    """Mirror of arn_test_iter_t (include/arnerf.h)."""
    _fields_ = [("rays_o", P), ("rays_d", P), ("hits_t", P), ("alive", P), ("n_alive", L),
                ("density_bitfield", P), ("cascades", I), ("grid_size", I), ("scale", F), ("exp_step_factor", F), ("n_samples", I), ("max_samples", I),
                ("T_threshold", F),
                ("xyz_min_host", P), ("xyz_max_host", P), ("levels", Levels), ("params_xyz_f16", P), ("params_rgb_f16", P), ("rgb_act", I),
                ("capacity", L), ("deltas", P), ("ts", P), ("n_eff", P), ("rays_a", P), ("counts", P), ("counts_alive", P),
                ("xyzs", P), ("dirs", P), ("sigmas", P), ("rgbs", P), ("ws", FieldWs),
                ("opacity", P), ("depth", P), ("rgb", P), ("alive_out", P), ("total_samples", P), ("schedule_rays", L)]


class SgTables(C.Structure):
    """Mirror of arn_sg_tables_t (include/arnerf.h)."""
    _fields_ = [("coeff_cl", P), ("D", I), ("H", I), ("W", I), ("C", I), ("components", P), ("mean", P), ("envH", I), ("envW", I),
                ("fh_tab", P), ("fh_h", I), ("fh_w", I), ("vol_range", F), ("angle_decay_fac", F), ("shadow_pow_fac", F),
                ("self_shadow_pow_fac", F)]


# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/arnerf.h one to one
SIGNATURES = {
    "arn_version": [],
    "arn_last_error": [],
    "arn_launch_count": [],
    "arn_set_tunable": [C.c_char_p, I],
    "arn_dbg_l2_red_peak": [P, L, L, P],
    "arn_p2p_alloc": [C.POINTER(C.c_void_p), L],
    "arn_p2p_free": [P],
    "arn_p2p_export": [P, C.c_char_p],
    "arn_p2p_open": [C.c_char_p, C.POINTER(C.c_void_p)],
    "arn_p2p_close": [P],
    "arn_p2p_signal": [C.POINTER(C.c_void_p), I, I, I, C.c_uint64, P],
    "arn_p2p_wait": [P, I, I, C.c_uint64, P],
    "arn_p2p_barrier": [C.POINTER(C.c_void_p), P, I, I, I, C.c_uint64, P],
    "arn_p2p_adam_exchange": [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), I, L, L, P, P, P, F, F, F, F, I, F, P],
    "arn_p2p_adam_exchange_mc": [P, P, L, L, P, P, P, F, F, F, F, I, F, P],
    "arn_p2p_set_timeout": [C.c_double],
    "arn_p2p_set_error_word": [P],
    "arn_p2p_set_grid": [I],
    "arn_gather_rays": [P, P, I, P, P, L, P, L, P, P, P],
    "arn_gather_batch": [P, P, I, P, P, L, P, L, P, L, I, P, P, P, P],
    "arn_grid_cell_positions": [P, P, L, I, F, P, P],
    "arn_grid_sample_cells": [P, F, I, F, P, P, L, P, P, P, P, P],
    "arn_grid_sample_cells_sorted": [P, F, I, F, P, P, L, P, P, P, P, P],
    "arn_grid_scatter": [P, L, P, P, L, P],
    "arn_density_grid_update": [P, P, P, F, F, L, P, P, P],
    "arn_mark_invisible_cells": [P, P, L, I, F, P, I, P, F, F, F, P, P, P],
    "arn_profile_enable": [I],
    "arn_profile_report": [C.c_char_p, I],
    "arn_ray_aabb_intersect": [P, P, L, P, P, I, I, P, P, P, P],
    "arn_ray_sphere_intersect": [P, P, L, P, P, I, I, P, P, P, P],
    "arn_ray_aabb_near": [P, P, L, P, P, F, P, P],
    "arn_morton3d": [P, L, P, P],
    "arn_morton3d_invert": [P, L, P, P],
    "arn_packbits": [P, I, F, P, L, P],
    "arn_march_train_count": [P, P, P, L, P, I, I, F, F, P, I, P, P, P],
    "arn_march_train_emit": [P, P, P, L, P, I, I, F, F, P, I, P, P, P, P, P, L, P],
    "arn_march_train_count_ex": [P, P, P, L, P, I, I, F, F, P, I, P, P, P, P, P],
    "arn_march_train_emit_ex": [P, P, P, L, P, I, I, F, F, P, I, P, P, P, P, P, P, L, P],
    "arn_march_test": [P, P, P, P, L, P, I, I, F, F, I, I, P, P, P, P, P, P],
    "arn_composite_train_fw": [P, P, P, P, P, L, L, F, P, P, P, P, P, P],
    "arn_composite_train_bw": [P, P, P, P, P, P, P, P, P, P, P, P, P, L, L, F, P, P, P],
    "arn_composite_test_fw": [P, P, P, P, P, L, I, F, P, P, P, P, P],
    "arn_distortion_fw": [P, P, P, P, L, L, P, P, P, P],
    "arn_distortion_bw": [P, P, P, P, P, P, P, L, L, P, P],
    "arn_march_train_bw": [P, P, P, P, L, P, P, P],
    "arn_hashgrid_geometry": [I, I, F, I, P, P, P, P],
    "arn_cast_f32_to_f16": [P, P, L, P],
    "arn_field_fw": [P, P, L, P, P, Levels, P, P, I, FieldWs, P, P, P],
    "arn_field_bw": [P, L, P, P, Levels, P, P, I, FieldWs, P, P, P, P, F, P, P, P, P, P],
    "arn_field_fw_simt": [P, P, L, P, P, Levels, P, P, I, FieldWs, P, P, P],
    "arn_field_fw_tc": [P, P, L, P, P, Levels, P, P, I, FieldWs, P, P, P],
    "arn_field_bw_tc": [P, L, P, P, Levels, P, P, I, FieldWs, P, P, P, P, F, P, P, P, P, P],
    "arn_field_fw_tc_dyn": [P, P, L, P, P, P, Levels, P, P, I, FieldWs, P, P, P],
    "arn_field_bw_tc_dyn": [P, L, P, P, P, Levels, P, P, I, FieldWs, P, P, P, P, F, P, P, P, P, P],
    "arn_hash_encode_fw_dyn": [P, L, P, P, P, Levels, P, P, P],
    "arn_hash_encode_bw_dyn": [P, L, P, P, P, Levels, P, P, P, P, P],
    "arn_march_train_emit_dyn": [P, P, L, I, I, F, F, I, P, P, P, P, P, P, P, L, P],
    "arn_nerf_loss": [P, P, P, P, L, P, F, F, F, F, P, P, P, P, P, P],
    "arn_composite_train_fw_loss": [P, P, P, P, P, L, L, F, P, P, P, P, P, P, P, F, F, F, F, P, P, P, P, P, P],
    "arn_train_fwbw": [C.POINTER(TrainCfg), P],
    "arn_render_test_iter": [C.POINTER(TestIterCfg), P],
    "arn_march_test_far_clamp": [P, P, P, L, P, I, I, F, F, I, P],
    "arn_march_test_all": [P, P, P, L, P, I, I, F, F, I, I, P, P, P, P],
    "arn_render_test_step_pre": [C.POINTER(TestIterCfg), P, P, P, P, P, P, I, I, L, P],
    "arn_render_test_step_fused": [C.POINTER(TestIterCfg), P, P, P, P, P, P, I, I, L, P],
    "arn_render_test_step": [C.POINTER(TestIterCfg), P, P, P, I, I, L, P],
    "arn_train_march": [C.POINTER(TrainCfg), P],
    "arn_train_set_fork": [I, P],
    "arn_train_set_join": [I, P],
    "arn_train_set_level_groups": [I, C.POINTER(C.c_int), C.POINTER(C.c_void_p)],
    "arn_train_fwbw_marched": [C.POINTER(TrainCfg), P],
    "arn_field_bw_simt": [P, L, P, P, Levels, P, P, I, FieldWs, P, P, P, P, F, P, P, P, P, P],
    "arn_hash_encode_fw": [P, L, P, P, Levels, P, P, P],
    "arn_hash_encode_bw": [P, L, P, P, Levels, P, P, P, P, P],
    "arn_sh4": [P, L, P, P],
    "arn_sg_shadow_factor": [C.POINTER(SgTables), P, I, P, L, P, P, F, P, P, P],
    "arn_sg_shade": [C.POINTER(SgTables), P, P, I, P, L, P, P, F, P, P, P, P, P, I, I, P, P, P, P],
    "arn_sg_shade_px": [P, I, I, L, P, P, P, P, P, I, P, P],
    "arn_sf_soft_shadow": [P, I, I, I, I, F, P, P, L, P, P, F, P, P, P],
    "arn_adam_step": [P, P, P, P, P, L, F, F, F, F, I, F, I, P],
    "arn_adam_step2": [P, P, P, P, P, L, P, P, P, P, P, L, F, F, F, F, I, F, I, P],
}
_RESTYPES = {"arn_last_error": C.c_char_p, "arn_launch_count": C.c_int64}

_lib = None


def lib():
    """Load libarnerf.so once.  Raises RuntimeError when it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C ar_nerf_b200/csrc` (or __graft_entry__.build()). "
                "ar_nerf_b200 has no CPU / PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = l
        # A/B switches without code changes: ARN_TUNABLES="hash_bw_blocks=2,march_warp=0" (arn_set_tunable names)
        for kv in filter(None, os.environ.get("ARN_TUNABLES", "").split(",")):
            name, _, val = kv.partition("=")
            if l.arn_set_tunable(name.strip().encode(), int(val)) != 0:
                raise RuntimeError(f"ARN_TUNABLES: {l.arn_last_error().decode()}")
    return _lib


# Optional per-entry-point device timing (bench.py's instrumented pass): when TIMING is a dict, every call is
# bracketed by CUDA events on the launching stream and (start, end) pairs are appended under the entry-point name.
TIMING = None


def call(name, *args):
    """Call an int-returning entry point; raise RuntimeError(arn_last_error()) on failure."""
    l = lib()
    if TIMING is not None and torch.cuda.is_available():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(l, name)(*args)
        e1.record()
        TIMING.setdefault(name, []).append((e0, e1))
    else:
        rc = getattr(l, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {l.arn_last_error().decode()}")


def timing_summary():
    """{entry point: (calls, total ms)} for the pairs collected since TIMING was set; synchronises."""
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (TIMING or {}).items()}


def launch_count():
    return int(lib().arn_launch_count())


def set_tunable(name, value):
    """Select a kernel variant (include/arnerf.h: arn_set_tunable); all variants compute the same results."""
    call("arn_set_tunable", name.encode(), int(value))


def profile_enable(on=True):
    call("arn_profile_enable", 1 if on else 0)


def profile_report():
    """{kernel: (calls, total device ms)} since the last report (CUDA events inside libarnerf.so)."""
    buf = C.create_string_buffer(1 << 16)
    call("arn_profile_report", buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, calls, ms = line.split()
        out[name] = (int(calls), float(ms))
    return out


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def check_tensor(t, name, dtype=None, ndim=None, last=None):
    """Reference behaviour (include/utils.h:4-6: CUDA + contiguous) plus dtype/shape validation."""
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError(f"{name} must have {ndim} dims, got shape {tuple(t.shape)}")
    if last is not None and t.shape[-1] != last:
        raise RuntimeError(f"{name} must have last dim {last}, got shape {tuple(t.shape)}")
    return t
