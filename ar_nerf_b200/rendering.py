"""Render API of the reference (models/rendering.py:13-54,162-319), kept verbatim at the call boundary:
`render(model, rays_o, rays_d, **kwargs) -> dict` with the same kwargs and result keys, so train.py:84-109,
show_gui.py:94 and insert/main.py:126,645 can call it unchanged.  The kernels underneath are libarnerf.so."""
import torch
import torch.nn.functional as F

from . import vren
from .custom_functions import RayAABBIntersector, RayMarcher, VolumeRenderer

MAX_SAMPLES = 1024    # rendering.py:9
NEAR_DISTANCE = 0.01  # rendering.py:10


# ---- the two helpers rendering.py pulls from insert/insert_utils.py (:18-19, :102-147) without open3d/matplotlib (Q13)
def normalize_eps(vec, eps=1e-6):
    return vec / (torch.norm(vec, dim=-1, keepdim=True) + eps)


def get_SH_val(shec, dirs, clamp_postive=False):
    """insert_utils.py:142-147 with SH_functions_torch (:102-131): order-3 real SH (9 terms), shec (9,3) -> (x,3)."""
    x, y, z = dirs[..., 0], dirs[..., 1], dirs[..., 2]
    dir_shs = torch.stack([
        0.2820947918 * torch.ones_like(x),
        0.4886025119 * y, 0.4886025119 * z, 0.4886025119 * x,
        1.0925484306 * x * y, 1.0925484306 * y * z, 0.3153915653 * (3.0 * z ** 2 - 1),
        1.0925484306 * x * z, 0.5462742153 * (x ** 2 - y ** 2)], dim=-1)
    vals = torch.matmul(dir_shs.unsqueeze(-2), shec).squeeze(-2)
    if clamp_postive:
        vals = F.relu(vals)
    return vals


def _intersect(model, rays_o, rays_d):
    """rendering.py:29-31.  Single scene box: fused slab test + near clamp; otherwise the general path."""
    if model.center.shape[0] == 1 and hasattr(model, 'host_box'):
        center, half_size = model.host_box()
        return vren.ray_aabb_near(rays_o.float(), rays_d.float(), center, half_size, NEAR_DISTANCE)
    _, hits_t, _ = RayAABBIntersector.apply(rays_o, rays_d, model.center, model.half_size, 1)
    hits_t[(hits_t[:, 0, 0] >= 0) & (hits_t[:, 0, 0] < NEAR_DISTANCE), 0, 0] = NEAR_DISTANCE
    return hits_t


@torch.amp.autocast('cuda')
def render(model, rays_o, rays_d, **kwargs):
    """
    Render rays by
    1. Compute the intersection of the rays with the scene bounding box
    2. Follow the process in @render_func (different for train/test)

    Inputs:
        model: NGP
        rays_o: (N_rays, 3) ray origins
        rays_d: (N_rays, 3) ray directions

    Outputs:
        result: dictionary containing final rgb and depth
    """
    rays_o = rays_o.contiguous(); rays_d = rays_d.contiguous()
    hits_t = _intersect(model, rays_o, rays_d)

    if kwargs.get('test_time', False):
        render_func = __render_rays_test
    else:
        render_func = __render_rays_train

    mesh_depth_map = kwargs.get('mesh_depth_map', None)
    if mesh_depth_map is not None:  # flattened (rendering.py:38-44)
        valid_depth = mesh_depth_map >= 1e-6
        hits_t_s = hits_t[valid_depth]
        update_min = torch.min(hits_t_s[:, 0, 1], mesh_depth_map[valid_depth])
        update_min = torch.max(update_min, hits_t_s[:, 0, 0])
        hits_t[valid_depth, 0, 1] = update_min

    results = render_func(model, rays_o, rays_d, hits_t, **kwargs)
    for k, v in results.items():
        if kwargs.get('to_cpu', False):
            v = v.cpu()
            if kwargs.get('to_numpy', False):
                v = v.numpy()
        results[k] = v
    return results


_TEST_WS = {}
_BOOST_SAMPLES = 2_560_000  # samples per iteration the boosted schedule aims at (see _render_rays_test_fused)
_STATE_RING = 4   # read-back slots of the device-driven test loop
_STATE_LAG = 2    # the host looks at the control state of the iteration queued this many calls earlier
_PREMARCH_MAX_BYTES = 12 << 30  # largest per-frame sample table (stride x N_rays floats) the fused test loop allocates
_GRAPH_ITERS = 8  # loop iterations per CUDA-graph replay (even: the ping-pong buffers are back in place after a replay)


def release_test_workspace():
    """Frees the buffers the fused test loop keeps between frames (sample table, per-ray state, captured graphs)."""
    _TEST_WS.clear()


def _make_test_ws(R, min_samples, device, boost=1):
    """Buffers of one device-driven test loop over R rays: every iteration marches at most n_alive * N_samples <=
    R * max(min_samples, boost) samples (N_samples = max(min(boost * R // n_alive, 64), min_samples))."""
    from .field import tile_rows, _scratch
    cap = R * max(min_samples, boost)
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)
    wimg = _scratch(device)
    return dict(cap=cap, deltas=f(cap), ts=f(cap), n_eff=torch.empty(R, dtype=torch.int32, device=device),
                rays_a=torch.empty(R, 3, dtype=torch.int64, device=device), counts=torch.zeros(4, dtype=torch.int32, device=device),
                xyzs=f(cap, 3), dirs=f(cap, 3), sigmas=f(cap), rgbs=f(cap, 3),
                feat=torch.empty(tile_rows(cap), 32, dtype=torch.float16, device=device), wimg=wimg,
                alive=[torch.empty(R, dtype=torch.int64, device=device), torch.empty(R, dtype=torch.int64, device=device)],
                total=torch.zeros(1, dtype=torch.int64, device=device),
                # device-driven loop (arn_render_test_step): two control-state buffers, chunk sums, pinned read-back ring
                state=torch.zeros(2, 8, dtype=torch.int32, device=device), partial=torch.empty((R + 127) // 128, dtype=torch.int32, device=device),
                state_host=torch.zeros(_STATE_RING, 8, dtype=torch.int32).pin_memory(),
                state_init=torch.zeros(8, dtype=torch.int32).pin_memory(),
                state_ev=[torch.cuda.Event() for _ in range(_STATE_RING)],
                # graph-driven loop: the frame's inputs / outputs live at fixed addresses
                g_rays_o=f(R, 3), g_rays_d=f(R, 3), g_hits=f(R, 2), g_opacity=f(R), g_depth=f(R), g_rgb=f(R, 3),
                g_arange=torch.arange(R, dtype=torch.int64, device=device), g_state0=torch.zeros(8, dtype=torch.int32, device=device),
                sync=torch.zeros(4, dtype=torch.int32, device=device), graphs={}, replays_hint=1, n_rays=R)


def _test_workspace(R, min_samples, device, boost=1):
    """The fused test loop's buffers, kept per (rays, min_samples, samples_boost, device) -- one frame size at a time."""
    key = (R, min_samples, boost, device.type, device.index)
    ws = _TEST_WS.get(key)
    if ws is None:
        ws = _make_test_ws(R, min_samples, device, boost)
        _TEST_WS.clear()
        _TEST_WS[key] = ws
    return ws


def _premarch_table(w, stride, N_rays, device):
    pm = w.get('premarch')
    if pm is None or pm[0].numel() < stride * N_rays:
        try:
            pm = (torch.empty(stride * N_rays, dtype=torch.float32, device=device), torch.empty(N_rays, dtype=torch.int32, device=device),
                  torch.empty(N_rays, dtype=torch.int32, device=device))
        except torch.cuda.OutOfMemoryError:  # no room for the sample table next to whatever else lives on the GPU: march per iteration
            pm = None
        w['premarch'] = pm
        w['graphs'].clear()  # captured graphs hold the old table's address
    return pm


def _test_cfg(model, w, rays_o, rays_d, hits_t2, opacity, depth, rgb, exp_step_factor, T_threshold, schedule_rays=0):
    """arn_test_iter_t over the given frame buffers (alive lists / n_alive are filled in per call)."""
    import ctypes as C
    from ._lib import FieldWs, TestIterCfg, ptr
    st = model.field_state
    p16x = st.cache_xyz.get(model.xyz_encoder.params); p16c = st.cache_rgb.get(model.rgb_net.params)
    cast = lambda a: C.cast(a, C.c_void_p)
    return TestIterCfg(
        ptr(rays_o), ptr(rays_d), ptr(hits_t2), None, 0,
        ptr(model.density_bitfield), model.cascades, model.grid_size, float(model.scale), float(exp_step_factor), 1, MAX_SAMPLES,
        float(T_threshold),
        cast(st.mn), cast(st.mx), st.geometry.c_levels, ptr(p16x), ptr(p16c), st.rgb_act,
        w['cap'], ptr(w['deltas']), ptr(w['ts']), ptr(w['n_eff']), ptr(w['rays_a']), w['counts'].data_ptr(), w['counts'].data_ptr() + 8,
        ptr(w['xyzs']), ptr(w['dirs']), ptr(w['sigmas']), ptr(w['rgbs']),
        FieldWs(ptr(w['feat']), None, None, None, None, None, ptr(w['wimg'])),
        ptr(opacity), ptr(depth), ptr(rgb), None, ptr(w['total']), int(schedule_rays)), (p16x, p16c)


def _render_test_graph(model, w, rays_o, rays_d, hits_t2, exp_step_factor, T_threshold, max_samples, min_samples, stride, launches=7, boost=1):
    """The device-driven loop replayed from CUDA graphs: ONE graph launch covers the frame's prologue (march of every ray,
    state / alive-list / output initialisation) plus the first _GRAPH_ITERS iterations, every further launch _GRAPH_ITERS more
    (an iteration behind the loop's end is a handful of empty kernels).  The frame's inputs are copied to fixed addresses;
    the host reads the control state once per batch of replays -- as many as the previous frame of this size needed.
    (Measured and dropped: dealing a small frame's rays to 2-8 independent loops captured as parallel graph branches -- 1.94 ms
    as one loop, 2.10 / 2.25 / 2.71 ms as 2 / 4 / 8 for one rank's eighth of an 800x800 frame.)"""
    import ctypes as C
    from ._lib import call, ptr, stream
    N_rays, device = rays_o.shape[0], rays_o.device
    ts_all, totals, cursor = w['premarch']
    st = model.field_state
    key = (ptr(model.density_bitfield), model.cascades, model.grid_size, float(model.scale), float(exp_step_factor), float(T_threshold), int(max_samples),
           int(min_samples), int(stride), st.cache_xyz.get(model.xyz_encoder.params).data_ptr(), st.cache_rgb.get(model.rgb_net.params).data_ptr(),
           tuple(st.mn), tuple(st.mx), id(st.geometry), st.rgb_act, launches, ts_all.data_ptr(), boost)
    graphs = w['graphs'].get(key)
    w['g_rays_o'].copy_(rays_o); w['g_rays_d'].copy_(rays_d); w['g_hits'].copy_(hits_t2)
    if graphs is None:
        cfg, keep = _test_cfg(model, w, w['g_rays_o'], w['g_rays_d'], w['g_hits'], w['g_opacity'], w['g_depth'], w['g_rgb'], exp_step_factor, T_threshold,
                              schedule_rays=boost * N_rays)
        cfg.n_alive = N_rays
        state_ptr = (w['state'][0].data_ptr(), w['state'][1].data_ptr())
        S0 = max(min(boost, 64), min_samples)
        w['g_state0'].copy_(torch.tensor([N_rays, S0, S0, 1 if max_samples > 0 else 0, 0, 0, 0, 0], dtype=torch.int32))

        def prologue():
            call("arn_march_test_all", ptr(w['g_rays_o']), ptr(w['g_rays_d']), ptr(w['g_hits']), N_rays, ptr(model.density_bitfield), model.cascades,
                 model.grid_size, float(model.scale), float(exp_step_factor), MAX_SAMPLES, stride, ptr(ts_all), ptr(totals), ptr(cursor), stream())
            w['alive'][0].copy_(w['g_arange']); w['state'][0].copy_(w['g_state0'])
            w['total'].zero_(); w['g_opacity'].zero_(); w['g_depth'].zero_(); w['g_rgb'].zero_(); w['sync'].zero_()

        def iterations():
            # launches = 7: arn_render_test_step_pre (the default: ordered lists, the two scans as kernels of their own); 4:
            # arn_render_test_step_fused (slice + emit | hash grid | MLP | compositing + survivors + state) -- same pixels, same
            # counts, lists in arrival order.  Measured equal (800x800: 5.48 / 5.52 ms; an eighth of it: 2.00 / 1.95 ms): inside a
            # graph the three small kernels cost next to nothing, an iteration is the latency of its hash-grid and MLP launches
            for it in range(_GRAPH_ITERS):
                cfg.alive, cfg.alive_out = w['alive'][it & 1].data_ptr(), w['alive'][(it & 1) ^ 1].data_ptr()
                if launches == 4:
                    call("arn_render_test_step_fused", C.byref(cfg), state_ptr[it & 1], state_ptr[(it & 1) ^ 1], ptr(w['sync']), ptr(ts_all), ptr(totals),
                         ptr(cursor), min_samples, int(max_samples), N_rays, stream())
                else:
                    call("arn_render_test_step_pre", C.byref(cfg), state_ptr[it & 1], state_ptr[(it & 1) ^ 1], ptr(w['partial']), ptr(ts_all), ptr(totals),
                         ptr(cursor), min_samples, int(max_samples), N_rays, stream())

        # one eager pass first: module loading and the kernels' one-time attribute calls are not capturable
        prologue(); iterations()
        torch.cuda.current_stream().synchronize()
        first, more = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(first, capture_error_mode="thread_local"):
            prologue(); iterations()
        with torch.cuda.graph(more, capture_error_mode="thread_local"):
            iterations()
        graphs = (first, more, keep)
        w['graphs'].clear()  # one configuration at a time
        w['graphs'][key] = graphs
    first, more, _ = graphs
    first.replay()
    queued, want = 1, max(1, w['replays_hint'])
    while True:
        while queued < want:
            more.replay(); queued += 1
        w['state_host'][0].copy_(w['state'][0], non_blocking=True)   # _GRAPH_ITERS is even: the last state written is state[0]
        w['state_ev'][0].record(torch.cuda.current_stream())
        w['state_ev'][0].synchronize()
        active, live_iters = int(w['state_host'][0][3]), int(w['state_host'][0][5])
        if not active or queued * _GRAPH_ITERS > int(max_samples):  # every live iteration requests at least one sample
            break
        want = queued + 1
    w['replays_hint'] = max(1, -(-live_iters // _GRAPH_ITERS))
    return w['g_opacity'].clone(), w['g_depth'].clone(), w['g_rgb'].clone(), w['total'][0].clone()


@torch.no_grad()
def _render_rays_test_fused(model, rays_o, rays_d, hits_t, **kwargs):
    """The loop of rendering.py:162-236 with each iteration as ONE native call (arn_render_test_iter: march, compact sample
    list, field, compositing + ray kill, alive-list compaction) and one host read per iteration -- the two counts that
    drive the reference's schedule.  Same schedule, same per-ray arithmetic, same results."""
    import ctypes as C
    from . import _lib
    from ._lib import call, ptr, stream
    exp_step_factor = kwargs.get('exp_step_factor', 0.)
    T_threshold = kwargs.get('T_threshold', 1e-4)
    max_samples = kwargs.get('max_samples', MAX_SAMPLES)
    N_rays, device = len(rays_o), rays_o.device
    min_samples = 1 if exp_step_factor == 0 else 4
    host_driven = kwargs.get('host_driven_test_loop', False)
    # samples_boost = k: the device-driven loop takes N_samples = max(min(k * N_rays // N_alive, 64), min_samples) per iteration
    # instead of the reference's k = 1 (whose N_rays samples per iteration bound ITS memory): fewer, larger iterations for a
    # loop that is bound by the NUMBER of its iterations (800x800: 41 -> 13 at k = 4, 6 at k = 16).  A ray's samples do not
    # depend on the slicing; its compositing does in the last bit (composite_test_fw restarts each iteration from
    # T = 1 - opacity instead of the running product: measured max |d rgb| 1.2e-7), and a ray that ends inside a slice has the
    # rest of that slice evaluated for nothing (total_samples counts them, as the reference's does).  k = 1 is the
    # reference's slicing bit for bit.  Default: k = 1 wherever the sample budget can bind (max_samples below the march's own
    # 1024, or exp_step_factor > 0: WHICH samples a ray gets then depends on the schedule); otherwise (the reference's
    # bounded-scene test frame) k = what keeps an iteration's buffers inside L2, ~2.5 M samples, at most 16.
    boost = kwargs.get('samples_boost')
    if boost is None:
        boost = min(16, _BOOST_SAMPLES // max(1, N_rays)) if (exp_step_factor == 0 and max_samples >= MAX_SAMPLES) else 1
    boost = 1 if host_driven else max(1, min(64, int(boost)))
    w = _test_workspace(N_rays, min_samples, device, boost)
    model.host_box()
    hits_t2 = hits_t[:, 0]
    if not hits_t2.is_contiguous():
        raise RuntimeError("hits_t must be contiguous")
    rays_o = rays_o.contiguous().float(); rays_d = rays_d.contiguous().float()
    s_ = stream()
    # The frame's samples are marched once, in front of the loop (arn_march_test_all), when their table fits: the loop never
    # asks a ray for more than max_samples + 63 samples.  Otherwise the iterations march (far-clamped rays).
    stride = int(max(1, max_samples)) + 64
    premarch = (not host_driven and kwargs.get('premarch_test_loop', True) and max_samples > 0 and stride * N_rays * 4 <= _PREMARCH_MAX_BYTES
                and _premarch_table(w, stride, N_rays, device) is not None)
    if premarch and kwargs.get('graph_test_loop', True) and not torch.cuda.is_current_stream_capturing():
        return _render_test_graph(model, w, rays_o, rays_d, hits_t2, exp_step_factor, T_threshold, max_samples, min_samples, stride,
                                  launches=4 if kwargs.get('test_loop_launches', 7) == 4 else 7, boost=boost)
    opacity = torch.zeros(N_rays, device=device)
    depth = torch.zeros(N_rays, device=device)
    rgb = torch.zeros(N_rays, 3, device=device)
    cur = 0
    torch.arange(N_rays, out=w['alive'][0])
    w['total'].zero_()
    cfg, _keep = _test_cfg(model, w, rays_o, rays_d, hits_t2, opacity, depth, rgb, exp_step_factor, T_threshold, schedule_rays=boost * N_rays)
    if not host_driven:
        # Loop control on the device: iterations are queued without waiting for their counts; the control state comes back
        # through pinned memory and is looked at _STATE_LAG iterations late (an iteration queued after the loop has ended
        # is a handful of empty launches).  n_alive never grows, so a stale value still bounds the grids.
        if premarch:
            ts_all, totals, cursor = w['premarch']
            call("arn_march_test_all", ptr(rays_o), ptr(rays_d), ptr(hits_t2), N_rays, ptr(model.density_bitfield), model.cascades,
                 model.grid_size, float(model.scale), float(exp_step_factor), MAX_SAMPLES, stride, ptr(ts_all), ptr(totals), ptr(cursor), s_)
        elif kwargs.get('far_clamp', True):
            call("arn_march_test_far_clamp", ptr(rays_o), ptr(rays_d), ptr(hits_t2), N_rays, ptr(model.density_bitfield), model.cascades,
                 model.grid_size, float(model.scale), float(exp_step_factor), MAX_SAMPLES, s_)
        S0 = max(min(boost, 64), min_samples)
        w['state_init'][:5] = torch.tensor([N_rays, S0, S0, 1 if max_samples > 0 else 0, 0], dtype=torch.int32)
        w['state'][0].copy_(w['state_init'], non_blocking=True)
        cfg.n_alive = N_rays
        cfg.capacity = w['cap']
        n_upper, it = N_rays, 0
        state_ptr = (w['state'][0].data_ptr(), w['state'][1].data_ptr())
        cur_stream = torch.cuda.current_stream()
        while True:
            cfg.alive, cfg.alive_out = w['alive'][it & 1].data_ptr(), w['alive'][(it & 1) ^ 1].data_ptr()
            if premarch:
                call("arn_render_test_step_pre", C.byref(cfg), state_ptr[it & 1], state_ptr[(it & 1) ^ 1], ptr(w['partial']), ptr(ts_all), ptr(totals),
                     ptr(cursor), min_samples, int(max_samples), n_upper, s_)
            else:
                call("arn_render_test_step", C.byref(cfg), state_ptr[it & 1], state_ptr[(it & 1) ^ 1], ptr(w['partial']), min_samples,
                     int(max_samples), n_upper, s_)
            k = it % _STATE_RING
            w['state_host'][k].copy_(w['state'][(it & 1) ^ 1], non_blocking=True)
            w['state_ev'][k].record(cur_stream)
            it += 1
            if it >= _STATE_LAG:
                j = (it - _STATE_LAG) % _STATE_RING
                w['state_ev'][j].synchronize()
                n_next, _, _, active, _ = w['state_host'][j][:5].tolist()
                if not active:
                    break
                n_upper = n_next
        return opacity, depth, rgb, w['total'][0].clone()
    samples, N_alive = 0, N_rays
    while samples < max_samples and N_alive > 0:
        N_samples = max(min(N_rays // N_alive, 64), min_samples)
        samples += N_samples
        cfg.alive, cfg.alive_out = w['alive'][cur].data_ptr(), w['alive'][cur ^ 1].data_ptr()
        cfg.n_alive, cfg.n_samples = N_alive, N_samples
        cfg.capacity = N_alive * N_samples  # this iteration's bound (<= the workspace's): small iterations skip the pipelining
        call("arn_render_test_iter", C.byref(cfg), s_)
        n_valid, _, n_keep, _ = w['counts'].tolist()  # the one host read of the iteration
        if n_valid == 0:
            break  # rendering.py:206 (nothing was composited, nothing changed)
        cur ^= 1
        N_alive = n_keep
    return opacity, depth, rgb, w['total'][0].clone()


@torch.no_grad()
def __render_rays_test(model, rays_o, rays_d, hits_t, **kwargs):
    """rendering.py:162-253: iterative march / evaluate / composite with alive-ray compaction (schedule kept, Q10)."""
    exp_step_factor = kwargs.get('exp_step_factor', 0.)
    results = {}

    N_rays = len(rays_o)
    device = rays_o.device
    fused = (getattr(model, 'field_impl', '') == '' and model.rgb_act == 'Sigmoid' and not model.use_raw_HDR
             and not kwargs.get('eager_test_loop', False) and N_rays > 0)
    # val_batch_size (rendering.py:209-215) only bounds the reference's per-call field batch: the field is a per-sample
    # function, so chunking cannot change a result; the fused loop's workspace holds a whole iteration and ignores it
    if fused:
        opacity, depth, rgb, total_samples = _render_rays_test_fused(model, rays_o, rays_d, hits_t, **kwargs)
        return _finish_test(results, opacity, depth, rgb, total_samples, rays_d, **kwargs)
    # ---- the reference's loop op by op (the cross-check of the fused loop, and the path of the HDR / CUDA-core field variants)
    opacity, depth, rgb = (torch.zeros(N_rays, *tail, device=device) for tail in ((), (), (3,)))
    alive = torch.arange(N_rays, device=device)
    hits_near_far = hits_t[:, 0]                       # (R,2) view, marched in place
    budget, threshold = kwargs.get('max_samples', MAX_SAMPLES), kwargs.get('T_threshold', 1e-4)
    fewest = 1 if exp_step_factor == 0 else 4
    requested, total_samples = 0, 0
    while requested < budget and alive.numel() > 0:
        per_ray = max(min(N_rays // alive.numel(), 64), fewest)   # few alive rays -> many samples each (rendering.py:197-199)
        requested += per_ray
        xyzs, dirs, deltas, ts, n_eff = vren.raymarching_test(rays_o, rays_d, hits_near_far, alive, model.density_bitfield, model.cascades,
                                                             model.scale, exp_step_factor, model.grid_size, MAX_SAMPLES, per_ray)
        total_samples = total_samples + n_eff.sum()
        xyzs, dirs = xyzs.reshape(-1, 3), dirs.reshape(-1, 3)
        real = (dirs != 0).any(dim=1)                  # padding slots carry a zero direction
        if not bool(real.any()):
            break
        sigmas, rgbs = _field_on_real_samples(model, xyzs, dirs, real, kwargs)
        vren.composite_test_fw(sigmas.view(-1, per_ray), rgbs.view(-1, per_ray, 3), deltas, ts, hits_near_far, alive, threshold,
                               n_eff, opacity, depth, rgb)
        alive = alive[alive >= 0]                      # composite_test_fw marks the rays it has finished with -1
    return _finish_test(results, opacity, depth, rgb, total_samples, rays_d, **kwargs)


def _field_on_real_samples(model, xyzs, dirs, real, kwargs):
    """(sigmas (n), rgbs (n,3)) over the padded sample list: the field where `real`, zeros elsewhere; at most `val_batch_size`
    samples per field call (rendering.py:209-219)."""
    pts, views = xyzs[real], dirs[real]
    step = kwargs.get('val_batch_size', pts.shape[0])
    parts = [model(pts[i:i + step], views[i:i + step], **kwargs) for i in range(0, pts.shape[0], step)]
    sigmas, rgbs = xyzs.new_zeros(xyzs.shape[0]), xyzs.new_zeros(xyzs.shape[0], 3)
    sigmas[real] = torch.cat([p[0] for p in parts], 0)
    rgbs[real] = torch.cat([p[1] for p in parts], 0).float()
    return sigmas, rgbs


def _finish_test(results, opacity, depth, rgb, total_samples, rays_d, **kwargs):
    """rendering.py:238-253: the result dict; the background behind the rays -- an image (IM_bkg) before a spherical-harmonics
    environment (SH_bkg) before black -- is blended in unless blend_bkg=False."""
    background = kwargs.get('IM_bkg')
    if background is None:
        sh = kwargs.get('SH_bkg')
        background = torch.zeros(3, device=rgb.device) if sh is None else get_SH_val(sh, rays_d, clamp_postive=True)
    if kwargs.get('blend_bkg', True):
        rgb += background * (1 - opacity)[:, None]
    results.update(opacity=opacity, depth=depth, rgb=rgb, total_samples=total_samples)
    return results


def __render_rays_train(model, rays_o, rays_d, hits_t, **kwargs):
    """rendering.py:255-298: march -> field -> composite, then the background behind what is left of each ray (Q9)."""
    step_factor = kwargs.get('exp_step_factor', 0.)
    rays_a, xyzs, dirs, deltas, ts, marched = RayMarcher.apply(rays_o, rays_d, hits_t[:, 0], model.density_bitfield, model.cascades, model.scale,
                                                             step_factor, model.grid_size, MAX_SAMPLES, kwargs.get('noise', None))
    # per-ray tensors among the keyword arguments (exposure ...) follow their ray's samples (rendering.py:266-268)
    for name, value in kwargs.items():
        if torch.is_tensor(value) and name != 'noise':
            kwargs[name] = value[rays_a[:, 0]].repeat_interleave(rays_a[:, 2], 0)
    sigmas, rgbs = model(xyzs, dirs, **kwargs)
    composited, opacity, depth, rgb, ws = VolumeRenderer.apply(sigmas, rgbs.contiguous(), deltas, ts, rays_a, kwargs.get('T_threshold', 1e-4))
    if kwargs.get('random_bg', False):
        background = torch.rand(3, device=rays_o.device)
    else:  # white behind synthetic scenes (fixed step), black behind real ones
        background = torch.full((3,), 1.0 if step_factor == 0 else 0.0, device=rays_o.device)
    return {'deltas': deltas, 'ts': ts, 'rm_samples': marched, 'vr_samples': composited, 'opacity': opacity, 'depth': depth,
            'rgb': rgb + background * (1 - opacity)[:, None], 'ws': ws, 'rays_a': rays_a}


@torch.enable_grad()
def render_surface_normal(model, pts, **kwargs):
    """rendering.py:300-313: normals = -normalize(d sigma / d x)."""
    H, W, _ = pts.shape
    pts_grad = pts.reshape(-1, 3).detach().clone().requires_grad_(True)
    sigmas = model.density(pts_grad)
    normals = torch.autograd.grad(sigmas, pts_grad, torch.ones_like(sigmas))[0]
    normals = normals.reshape(H, W, 3).nan_to_num(0.0, 1.0, -1.0).detach()
    normals = -normalize_eps(normals)
    return normals


@torch.no_grad()
def render_surface_rgb(model, pts, rays_d, **kwargs):
    """rendering.py:315-319."""
    H, W, _ = pts.shape
    sigmas, rgbs = model(pts.reshape(-1, 3), rays_d.reshape(-1, 3), **kwargs)
    return rgbs.reshape(H, W, 3)
