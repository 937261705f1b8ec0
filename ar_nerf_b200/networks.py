"""NGP field of the reference (models/networks.py:12-281): same constructor, attributes, buffers, state-dict keys and
methods, evaluated by libarnerf.so (no tinycudann)."""
import numpy as np
import torch
import torch.nn.functional as F
from einops import rearrange
from torch import nn

from . import vren
from .custom_functions import TruncExp
from .field import Encoding, FieldFunction, FieldState, HashGeometry, Network, NetworkWithInputEncoding, field_inference

NEAR_DISTANCE = 0.01  # models/rendering.py:10 (imported from there by the reference; defined here to avoid the cycle)


class _Tonemapper(nn.Module):
    """tcnn.Network 1 -> 64 (ReLU) -> 1, Sigmoid (networks.py:80-93, HDR-NeRF mode only).  Secondary branch of the
    reference, kept as plain torch ops; params follow tiny-cuda-nn's layout (input padded to 16)."""

    def __init__(self, seed):
        super().__init__()
        gen = torch.Generator().manual_seed(seed)
        bound1, bound2 = (6.0 / (16 + 64)) ** 0.5, (6.0 / (64 + 16)) ** 0.5
        p = torch.cat([(torch.rand(64 * 16, generator=gen) * 2 - 1) * bound1, (torch.rand(16 * 64, generator=gen) * 2 - 1) * bound2])
        self.params = nn.Parameter(p)

    def forward(self, x):
        # tiny-cuda-nn pads the 1-wide input to the 16 columns of W1 with ONES (the padded columns act as a bias) and reads
        # output 0 of the 16-wide last layer; fp16 weights / activations, fp32 accumulate
        w1 = self.params[:1024].view(64, 16).half().float(); w2 = self.params[1024:].view(16, 64).half().float()
        xin = torch.cat([x.float(), torch.ones(x.shape[0], 15, dtype=torch.float32, device=x.device)], 1).half().float()
        hid = torch.relu(xin @ w1.T).half().float()
        return torch.sigmoid(hid @ w2[:1].T)


class NGP(nn.Module):
    def __init__(self, scale, rgb_act='Sigmoid', use_raw_HDR=False):
        super().__init__()
        self.rgb_act = rgb_act
        self.use_raw_HDR = use_raw_HDR

        # scene bounding box (networks.py:19-24)
        self.scale = scale
        self.register_buffer('center', torch.zeros(1, 3))
        self.register_buffer('xyz_min', -torch.ones(1, 3) * scale)
        self.register_buffer('xyz_max', torch.ones(1, 3) * scale)
        self.register_buffer('half_size', (self.xyz_max - self.xyz_min) / 2)

        # each density grid covers [-2^(k-1), 2^(k-1)]^3 for k in [0, C-1] (networks.py:26-30)
        self.cascades = max(1 + int(np.ceil(np.log2(2 * scale))), 1)
        self.grid_size = 128
        self.register_buffer('density_bitfield', torch.zeros(self.cascades * self.grid_size ** 3 // 8, dtype=torch.uint8))

        # networks.py:33-35
        L = 16; Fdim = 2; log2_T = 19; N_min = 16
        b = np.exp(np.log(2048 * scale / N_min) / (L - 1))
        print(f'GridEncoding: Nmin={N_min} b={b:.5f} F={Fdim} T=2^{log2_T} L={L}')

        self._per_level_scale_f64 = float(b)
        self.geometry = HashGeometry(L, N_min, float(np.float32(b)), log2_T)
        self.xyz_encoder = NetworkWithInputEncoding(self.geometry)
        self.dir_encoder = Encoding()
        self.rgb_net = Network()
        self.field_state = FieldState(self.geometry, self.xyz_min[0].tolist(), self.xyz_max[0].tolist(), rgb_act)
        self.field_impl = ""  # "" = production kernels, "_simt" = CUDA-core cross-check path (tests)

        if self.rgb_act == 'None' and not self.use_raw_HDR:  # rgb_net output is log-radiance (networks.py:80-93)
            for i in range(3):
                setattr(self, f'tonemapper_net_{i}', _Tonemapper(4242 + i))

    # -------------------------------------------------------------------------------------------- host copies
    def host_box(self):
        """(center, half_size) of the single scene box as python floats, cached (no device sync per render call)."""
        key = (self.center._version, self.half_size._version, self.center.data_ptr())
        if getattr(self, '_host_box_key', None) != key:
            self._host_box = (self.center[0].tolist(), self.half_size[0].tolist())
            self._host_box_key = key
            self.field_state.set_box(self.xyz_min[0].tolist(), self.xyz_max[0].tolist())
        return self._host_box

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # A checkpoint written by a tiny-cuda-nn build that sized the levels in double has a different table length
        # (5 710 032 instead of 5 722 520 entries at scale 0.5): adopt that geometry before loading.
        key = prefix + 'xyz_encoder.params'
        if key in state_dict and state_dict[key].numel() != self.xyz_encoder.params.numel():
            g = self.geometry
            alt = HashGeometry.in_double(g.n_levels, g.base_resolution, self._per_level_scale_f64, g.log2_hashmap_size)
            if state_dict[key].numel() == 3072 + 2 * alt.total:
                self.geometry = alt
                self.xyz_encoder.geometry = alt
                self.xyz_encoder.params = nn.Parameter(torch.empty(3072 + 2 * alt.total, dtype=self.xyz_encoder.params.dtype,
                                                                   device=self.xyz_encoder.params.device))
                self.field_state.geometry = alt
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._host_box_key = None

    # -------------------------------------------------------------------------------------------- field
    def density(self, x, return_feat=False):
        """networks.py:95-108.  x (N,3) in [-scale, scale] -> sigmas (N) [, h (N,16)]."""
        self.host_box()
        if self._needs_graph(x):
            sigmas, _, h = FieldFunction.apply(x, None, self.xyz_encoder.params, None, self.field_state, self.field_impl)
        else:  # no backward can follow: nothing is saved, no activation leaves the SM
            sigmas, _, h = field_inference(x, None, self.xyz_encoder.params, None, self.field_state, self.field_impl, want_h=return_feat)
        if return_feat:
            return sigmas, h
        return sigmas

    def _needs_graph(self, x):
        return torch.is_grad_enabled() and (x.requires_grad or self.xyz_encoder.params.requires_grad or self.rgb_net.params.requires_grad)

    def log_radiance_to_rgb(self, log_radiances, **kwargs):
        """networks.py:110-131."""
        log_exposure = torch.log(kwargs['exposure']) if 'exposure' in kwargs else 0
        out = []
        for i in range(3):
            inp = log_radiances[:, i:i + 1] + log_exposure
            out += [getattr(self, f'tonemapper_net_{i}')(inp)]
        return torch.cat(out, 1)

    def forward(self, x, d, **kwargs):
        """networks.py:133-165.  x (N,3), d (N,3) -> sigmas (N), rgbs (N,3)."""
        self.host_box()
        if self._needs_graph(x):
            sigmas, rgbs, _ = FieldFunction.apply(x, d, self.xyz_encoder.params, self.rgb_net.params, self.field_state,
                                                  self.field_impl)
        else:
            sigmas, rgbs, _ = field_inference(x, d, self.xyz_encoder.params, self.rgb_net.params, self.field_state, self.field_impl)
        if self.use_raw_HDR:
            rgbs = F.leaky_relu(rgbs) if not kwargs.get('output_radiance', False) else torch.relu(rgbs)
        elif self.rgb_act == 'None':
            if kwargs.get('output_radiance', False):
                rgbs = TruncExp.apply(torch.clamp(rgbs, 0, 20))
            else:
                rgbs = self.log_radiance_to_rgb(rgbs, **kwargs)
        return sigmas, rgbs

    # -------------------------------------------------------------------------------------------- occupancy grid
    def init_density_grid(self):
        """Registers `density_grid` (C,G^3) and `grid_coords` (G^3,3) exactly as train.py:79-82 does on the model."""
        G = self.grid_size
        dev = self.density_bitfield.device
        if not hasattr(self, 'density_grid'):
            self.register_buffer('density_grid', torch.zeros(self.cascades, G ** 3, device=dev))
        if not hasattr(self, 'grid_coords'):
            r = torch.arange(G, dtype=torch.int32, device=dev)
            zz, yy, xx = torch.meshgrid(r, r, r, indexing='ij')
            self.register_buffer('grid_coords', torch.stack([xx, yy, zz], -1).reshape(-1, 3).contiguous())

    @torch.no_grad()
    def get_all_cells(self):
        """networks.py:167-179.  grid_coords never changes, so its morton codes are computed once."""
        key = (self.grid_coords.data_ptr(), self.grid_coords._version)
        if getattr(self, '_all_cells_key', None) != key:
            self._all_cells_idx = vren.morton3D(self.grid_coords).long()
            self._all_cells_key = key
        return [(self._all_cells_idx, self.grid_coords)] * self.cascades

    @torch.no_grad()
    def sample_uniform_and_occupied_cells(self, M, density_threshold):
        """networks.py:181-207 (same torch RNG calls in the same order)."""
        cells = []
        for c in range(self.cascades):
            dev = self.density_grid.device
            coords1 = torch.randint(self.grid_size, (M, 3), dtype=torch.int32, device=dev)
            indices1 = vren.morton3D(coords1).long()
            # M occupied cells drawn uniformly with replacement (networks.py:196-201).  The reference materialises
            # nonzero(grid > thr) and indexes it with randint(len) -- a host round trip per cascade; here the k-th occupied
            # cell is found on the device through the running count of occupied cells: same distribution, no sync.
            occupied = self.density_grid[c] > density_threshold
            running = torch.cumsum(occupied, 0, dtype=torch.int32)
            count = running[-1]
            u = torch.randint(2 ** 31 - 1, (M,), device=dev)
            k = torch.remainder(u, count.clamp(min=1)).int()
            indices2 = torch.searchsorted(running, k + 1).clamp_(max=occupied.numel() - 1)
            # (an empty grid makes the reference skip this half; here it degenerates to re-sampling one cell, which is harmless)
            coords2 = vren.morton3D_invert(indices2.int())
            cells += [(torch.cat([indices1, indices2]), torch.cat([coords1, coords2]))]
        return cells

    @torch.no_grad()
    def mark_invisible_cells(self, K, poses, img_wh, chunk=64 ** 3):
        """networks.py:209-250: cells no camera covers, or that lie closer than NEAR_DISTANCE in front of a camera, get density
        -1 (never revived by update_density_grid); `count_grid` holds the covered fraction.  One native launch per cascade over
        all cells and cameras (arn_mark_invisible_cells); `chunk` is accepted for signature compatibility and unused."""
        self.count_grid = torch.zeros_like(self.density_grid)
        dev = self.density_grid.device
        c2w = poses.to(dev).float()
        rot = c2w[:, :3, :3].transpose(1, 2)                                   # world -> camera rotation
        w2c = torch.cat([rot.reshape(-1, 9), (-rot @ c2w[:, :3, 3:]).reshape(-1, 3)], 1).contiguous()
        for c, (indices, coords) in enumerate(self.get_all_cells()):
            vren.mark_invisible_cells(coords, indices, self.grid_size, min(2 ** (c - 1), self.scale), w2c, K, img_wh, NEAR_DISTANCE,
                                      self.density_grid[c], self.count_grid[c])
        self._grid_version = getattr(self, '_grid_version', 0) + 1  # a cell selection computed ahead of this call is stale

    # ---- steady-state cell selection, one refresh ahead.  Which cells a refresh samples depends on the density grid the PREVIOUS
    # refresh left behind and on random draws -- not on the weights -- so the draws, the selection (arn_grid_sample_cells_sorted:
    # cells in curve order, which makes the density evaluation's hash-grid gathers 1.8x cheaper) and the zero fill of the scratch
    # grid run on a side stream right behind a refresh, under the training steps that follow; the next refresh finds them done.
    def _grid_key(self, density_threshold):
        g = self.density_grid
        return (g.data_ptr(), g._version, getattr(self, '_grid_version', 0), float(density_threshold), self.grid_size, self.cascades, float(self.scale))

    def _refresh_buffers(self):
        G, dev = self.grid_size, self.density_grid.device
        M = G ** 3 // 4
        b = getattr(self, '_refresh_ws', None)
        if b is None or b['key'] != (G, self.cascades, dev):
            b = self._refresh_ws = dict(
                key=(G, self.cascades, dev), tmp=torch.zeros_like(self.density_grid), tmp_clean=True,
                cells=[(torch.empty(2 * M, dtype=torch.int64, device=dev), torch.empty(2 * M, 3, device=dev)) for _ in range(self.cascades)],
                # the side stream's selection has its own scratch: another model's refresh may run beside it
                scratch=torch.empty(vren.grid_sample_scratch_ints(G, M, sort=True), dtype=torch.int32, device=dev),
                side=torch.cuda.Stream(device=dev), ready=None, ready_key=None, gen=None, rng_delta=None)
        return b

    def _draw_and_select(self, b, density_threshold, sort, generator=None):
        """The two randint draws of sample_uniform_and_occupied_cells, cascade by cascade, in the reference's order, then the
        rand draw of the positions and the native selection (k-th occupied cell through bit masks) per cascade."""
        G, dev = self.grid_size, self.density_grid.device
        M = G ** 3 // 4
        draws = [(torch.randint(G, (M, 3), dtype=torch.int32, device=dev, generator=generator),
                  torch.randint(2 ** 31 - 1, (M,), device=dev, generator=generator)) for _ in range(self.cascades)]
        for c in range(self.cascades):
            rnd = torch.rand((2 * M, 3), dtype=torch.float32, device=dev, generator=generator)
            vren.grid_sample_cells(self.density_grid[c], density_threshold, G, min(2 ** (c - 1), self.scale), *draws[c], rnd, sort=sort, out=b['cells'][c],
                                   scratch=b['scratch'])

    def _prefetch_refresh(self, b, density_threshold):
        """Queue the NEXT refresh's draws + selection (+ the zero fill of the scratch grid) on the side stream.  Its draws come
        from a private generator seeded from the global seed; the refresh that uses them moves the global CUDA generator past the
        draws it would have made itself (b['rng_delta'], measured once on the global generator and rewound), so every OTHER draw
        of a seeded run -- the marcher's noise, a random background -- sits where it sits without the prefetch."""
        dev = self.density_grid.device
        glob = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
        if b['gen'] is None:
            b['gen'] = torch.Generator(device=dev)
            b['gen'].manual_seed((glob.initial_seed() + 0x9E3779B97F4A7C15) % (2 ** 63))
        if b['rng_delta'] is None:
            off = glob.get_offset()
            self._draw_and_select(b, density_threshold, sort=False)   # (the cells it selects are overwritten below)
            b['rng_delta'] = glob.get_offset() - off
            glob.set_offset(off)
        main = torch.cuda.current_stream()
        done = torch.cuda.Event(); done.record(main)       # this refresh has consumed the cells and the scratch grid
        with torch.cuda.stream(b['side']):
            b['side'].wait_event(done)
            b['tmp'].zero_(); b['tmp_clean'] = True
            self._draw_and_select(b, density_threshold, sort=True, generator=b['gen'])
            b['ready'] = torch.cuda.Event(); b['ready'].record(b['side'])
        b['ready_key'] = self._grid_key(density_threshold)

    @torch.no_grad()
    def update_density_grid(self, density_threshold, warmup=False, decay=0.95, erode=False, prefetch_next=None):
        """networks.py:252-281.  prefetch_next: whether the NEXT call will be a steady-state refresh (warmup=False) with the same
        threshold, so that its cell selection can be computed ahead (default: the kind of this call)."""
        G, dev = self.grid_size, self.density_grid.device
        native_sampling = G <= 160 and (G ** 3 // 4) % 256 == 0 and self.density_grid.is_cuda
        if not native_sampling:
            return self._update_density_grid_torch(density_threshold, warmup, decay, erode)
        b = self._refresh_buffers()
        main = torch.cuda.current_stream()
        if b['ready'] is not None:
            main.wait_event(b['ready'])   # (also orders the side stream's zero fill in front of the writes below)
        prefetched = b['ready'] is not None and not warmup and b['ready_key'] == self._grid_key(density_threshold)
        b['ready'] = None
        if not b['tmp_clean']:
            b['tmp'].zero_()
        if warmup:
            cells = self.get_all_cells()
        elif not prefetched:
            self._draw_and_select(b, density_threshold, sort=False)
        else:  # the global generator moves as if this refresh had drawn its cells here
            glob = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
            glob.set_offset(glob.get_offset() + b['rng_delta'])
        for c in range(self.cascades):
            if warmup:
                indices, coords = cells[c]
                # xyzs_w = (coords/(G-1)*2-1)*(s - s/G) + (rand*2-1)*(s/G): one kernel, same torch.rand_like draw, same bits
                rnd = torch.rand(coords.shape, dtype=torch.float32, device=coords.device)
                xyzs_w = vren.grid_cell_positions(coords, rnd, G, min(2 ** (c - 1), self.scale))
            else:
                indices, xyzs_w = b['cells'][c]
            vren.grid_scatter(b['tmp'][c], indices, self.density(xyzs_w))  # density_grid_tmp[c, indices] = density (:268)
        b['tmp_clean'] = False
        decay_cells = None
        if erode:
            decay_cells = torch.clamp(decay ** (1 / self.count_grid), 0.1, 0.95).float().contiguous()
        # where(grid < 0, grid, max(grid*decay, tmp)) in place, mean of the positive cells, threshold, packbits: no host sync
        vren.density_grid_update(self.density_grid, b['tmp'], decay_cells, decay, density_threshold, self.density_bitfield)
        self._grid_version = getattr(self, '_grid_version', 0) + 1
        if (not warmup) if prefetch_next is None else prefetch_next:
            self._prefetch_refresh(b, density_threshold)

    @torch.no_grad()
    def _update_density_grid_torch(self, density_threshold, warmup=False, decay=0.95, erode=False):
        """The same refresh where the native cell selection does not apply (grid_size > 160, odd sizes, CPU tensors)."""
        density_grid_tmp = torch.zeros_like(self.density_grid)
        cells = self.get_all_cells() if warmup else self.sample_uniform_and_occupied_cells(self.grid_size ** 3 // 4, density_threshold)
        for c in range(self.cascades):
            s = min(2 ** (c - 1), self.scale)
            indices, coords = cells[c]
            rnd = torch.rand(coords.shape, dtype=torch.float32, device=coords.device)
            xyzs_w = vren.grid_cell_positions(coords, rnd, self.grid_size, s)
            density_grid_tmp[c, indices] = self.density(xyzs_w)
        decay_cells = None
        if erode:
            decay_cells = torch.clamp(decay ** (1 / self.count_grid), 0.1, 0.95).float().contiguous()
        vren.density_grid_update(self.density_grid, density_grid_tmp, decay_cells, decay, density_threshold, self.density_bitfield)
