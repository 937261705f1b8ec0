"""Host-side multi-GPU logic: rays are independent, so training shards a global batch (or draws
per-rank batches) with no data-path collective except ONE gradient exchange per step; test rendering shards the frame
into contiguous row bands and gathers the per-ray results on rank 0."""
import os

import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous, balanced [lo, hi) of n items for `rank` of `world` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rays(rays_o, rays_d, rank, world, *extra):
    lo, hi = shard_bounds(rays_o.shape[0], rank, world)
    return tuple(t[lo:hi] for t in (rays_o, rays_d) + extra)


def shard_rays_interleaved(rays_o, rays_d, rank, world, *extra):
    """Rank r takes rays r, r + world, r + 2*world, ...: every rank sees the same mix of empty and occupied pixels, so the
    test-time render loop is balanced (contiguous bands give the middle ranks all the work of a centred object)."""
    return tuple(t[rank::world].contiguous() for t in (rays_o, rays_d) + extra)


def gather_frame_interleaved(local, n_total, rank, world, dst=0):
    """Inverse of shard_rays_interleaved for per-ray results (n_local, C): the (n_total, C) frame on `dst`, None elsewhere."""
    if world == 1:
        return local
    pad = (n_total + world - 1) // world
    send = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    send[:local.shape[0]] = local
    if dist.get_backend() == "nccl":
        out = torch.empty((world, pad) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out.view((world * pad,) + tuple(local.shape[1:])), send)
    else:
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send)
        out = torch.stack(parts, 0)
    if rank != dst:
        return None
    # ray i*world + r is row i of rank r; the padding rows all land behind n_total
    return out.transpose(0, 1).reshape((world * pad,) + tuple(local.shape[1:]))[:n_total]


def allreduce_grads(grads, world):
    """Sum-reduce gradient buffers in place (the optimizer divides by world); one collective per buffer."""
    if world > 1:
        for g in grads:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)


def gather_frame(local, n_total, rank, world, dst=0):
    """Gathers row-band results (n_local, C) of every rank into an (n_total, C) tensor on `dst` (None elsewhere)."""
    if world == 1:
        return local
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)  # collectives want equal shapes: pad every band to the largest one
    send = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    send[:local.shape[0]] = local
    if dist.get_backend() == "nccl":
        # one all-gather (a single NCCL kernel over NVSwitch) instead of gather's world-1 send/recv pairs; 12.8 MB per frame
        out = torch.empty((world * pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, send)
        if rank != dst:
            return None
        if all(hi - lo == pad for lo, hi in sizes):
            return out[:n_total]
        return torch.cat([out[r * pad:r * pad + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


# ---------------------------------------------------------------------------------------------------------------------
# Sharded optimizer state (ZeRO-1 style) for the one large parameter, the hash table: instead of all-reducing the 45.8 MB
# fp32 gradient and running the same 11.4 M-element Adam on every rank, the gradient is reduce-scattered, every rank
# updates ITS contiguous slice (fp32 master, moments, fp16 working copy) and the fp16 slices are all-gathered.
# Traffic per rank: (N-1)/N * (4 + 2) bytes per parameter instead of (N-1)/N * 8; Adam's HBM traffic drops by N.
def shard_size(n, world, align=8):
    """Elements per rank: ceil(n / world) rounded up to `align` (16-byte vector access of the fp16 copy / fp32 state)."""
    per = (n + world - 1) // world
    return (per + align - 1) // align * align


def padded_numel(n, world, align=8):
    return shard_size(n, world, align) * world


def reduce_scatter_sum(full, out, rank, world):
    """out (S) <- sum over ranks of full[rank*S:(rank+1)*S]; full has world*S elements."""
    if world == 1:
        out.copy_(full[:out.numel()])
        return
    if dist.get_backend() == "nccl":
        dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM)
    else:  # gloo has no reduce-scatter: same result through an all-reduce (CPU tests)
        tmp = full.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
        out.copy_(tmp[rank * out.numel():(rank + 1) * out.numel()])


def all_gather_shards(full, rank, world):
    """In place: every rank contributes full[rank*S:(rank+1)*S]; afterwards `full` is identical on all ranks."""
    if world == 1:
        return
    S = full.numel() // world
    mine = full[rank * S:(rank + 1) * S]
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(full, mine)
    else:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.clone())
        for r, part in enumerate(parts):
            full[r * S:(r + 1) * S].copy_(part)


# ---------------------------------------------------------------------------------------------------------------------
# Peer-memory exchange (csrc/arn_p2p.cu): the gradient buffer and the fp16 working copy of the sharded parameter live in
# cudaMalloc'ed memory whose CUDA IPC handles are exchanged once; afterwards ONE kernel per step reads the ranks'
# gradient slices over NVLink, runs Adam and writes the fp16 slice into every rank's copy.
class _RawCuda:
    """Zero-copy torch view of raw device memory (torch.as_tensor understands __cuda_array_interface__)."""

    def __init__(self, ptr, numel, typestr):
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def group_runs(bounds, world, rank):
    """Ownership of a parameter sharded GROUP BY GROUP: bounds = [0, b1, .., n] (multiples of 8).  Rank r owns, of group g,
    the run [b_g + r*S_g, b_g + (r+1)*S_g) clipped to the group (S_g = shard_size(len_g, world)).  Returns [(lo, cnt, off)]:
    start and length of the rank's run of each group in the flat parameter and where it starts in the rank's packed state."""
    runs, off = [], 0
    for a, b in zip(bounds, bounds[1:]):
        S = shard_size(b - a, world)
        lo = min(a + rank * S, b)
        cnt = max(0, min(lo + S, b) - lo)
        runs.append((lo, cnt, off))
        off += cnt
    return runs


class PeerExchange:
    """Buffers + peer mappings + flag slots of one sharded parameter.  Collective constructor (every rank calls it).

    bounds = element boundaries [0, b1, .., n] of GROUPS of the flat parameter (multiples of 8): every group is sharded over
    the ranks on its own -- rank r owns [b_g + r*S_g, b_g + (r+1)*S_g) of group g -- so that a group whose gradient is final
    can be exchanged while the hash-grid backward is still reducing the next one (step(..., group_events=)).  m / v hold the
    rank's owned elements group after group (`off` = where a group's run starts)."""

    def __init__(self, n, world, rank, device, bounds=None):
        import ctypes as C
        from . import _lib
        self.n, self.world, self.rank = n, world, rank
        bounds = [0, n] if bounds is None else [int(b) for b in bounds]
        assert bounds[0] == 0 and bounds[-1] == n and all(a < b and a % 8 == 0 for a, b in zip(bounds, bounds[1:])), bounds
        assert len(bounds) - 1 <= 16
        self.bounds = bounds
        self.groups = group_runs(bounds, world, rank)  # (lo, cnt, off): this rank's run of each group inside the flat parameter / inside m, v
        self.owned = sum(c for _, c, _ in self.groups)
        self.P = (n + 7) // 8 * 8

        def alloc(nbytes):
            out = C.c_void_p()
            _lib.call("arn_p2p_alloc", C.byref(out), int(nbytes))
            return out.value

        def export(ptr):
            buf = C.create_string_buffer(64)
            _lib.call("arn_p2p_export", ptr, buf)
            return buf.raw

        # Buffers every rank can address: torch symmetric memory when it rendezvouses (it also hands out MULTICAST addresses of
        # the ranks' buffers -- NVLS: the switch reduces and replicates, arn_p2p_adam_exchange_mc), else cudaMalloc + CUDA IPC
        # handles (peer loads / stores only).  ARN_P2P_BACKEND=ipc forces the latter.  Every rank must take the same branch.
        self.mc_grad = self.mc_p16 = None
        self._symm = None
        ok = torch.zeros(1, device=device)
        if os.environ.get("ARN_P2P_BACKEND", "symm") == "symm":
            try:
                import torch.distributed._symmetric_memory as symm
                gname = dist.group.WORLD.group_name
                with torch.cuda.device(device):
                    g = symm.empty(self.P, dtype=torch.float32, device=device)
                    h = symm.empty(self.P, dtype=torch.float16, device=device)
                    f = symm.empty(1024, dtype=torch.int32, device=device)
                    g.zero_(); h.zero_(); f.zero_()
                    hg, hh, hf = symm.rendezvous(g, gname), symm.rendezvous(h, gname), symm.rendezvous(f, gname)
                self._symm = (g, h, f, hg, hh, hf)
                ok.fill_(1.0)
            except Exception as e:  # noqa: BLE001 -- any failure means "use IPC"
                print(f"[ar_nerf_b200] symmetric memory unavailable on rank {rank} ({type(e).__name__}: {e}); using CUDA IPC")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 1.0:
            g, h, f, hg, hh, hf = self._symm
            self.grad, self.p16 = g, h
            self.g_ptr, self.h_ptr, self.f_ptr = g.data_ptr(), h.data_ptr(), f.data_ptr()
            self.G = (C.c_void_p * world)(*[int(x) for x in hg.buffer_ptrs])
            self.H = (C.c_void_p * world)(*[int(x) for x in hh.buffer_ptrs])
            self.F = (C.c_void_p * world)(*[int(x) for x in hf.buffer_ptrs])
            mg, mh = int(getattr(hg, "multicast_ptr", 0) or 0), int(getattr(hh, "multicast_ptr", 0) or 0)
            # the multicast form pays from 8 ranks on (0.430 -> 0.399 ms per step); at 4 it measures equal (0.407), at 2 the
            # switch's read of both copies costs more than the one peer load it replaces (0.379 -> 0.415): peer loads / stores there
            want_mc = os.environ.get("ARN_P2P_MULTICAST", "1" if world >= 8 else "0") == "1" and not os.environ.get("ARN_P2P_NO_MULTICAST")
            if mg and mh and want_mc:
                self.mc_grad, self.mc_p16 = mg, mh
        else:
            self._symm = None
            with torch.cuda.device(device):
                self.g_ptr, self.h_ptr, self.f_ptr = alloc(self.P * 4), alloc(self.P * 2), alloc(4096)
                mine = (export(self.g_ptr), export(self.h_ptr), export(self.f_ptr))
                everyone = [None] * world
                dist.all_gather_object(everyone, mine)
                arrs = []
                for k, own in enumerate((self.g_ptr, self.h_ptr, self.f_ptr)):
                    ptrs = []
                    for r in range(world):
                        if r == rank:
                            ptrs.append(own)
                        else:
                            out = C.c_void_p()
                            _lib.call("arn_p2p_open", everyone[r][k], C.byref(out))
                            ptrs.append(out.value)
                    arrs.append((C.c_void_p * world)(*ptrs))
                self.G, self.H, self.F = arrs
            self.grad = torch.as_tensor(_RawCuda(self.g_ptr, self.P, "<f4"), device=device)
            self.p16 = torch.as_tensor(_RawCuda(self.h_ptr, self.P, "<f2"), device=device)
        self.zero_stream = torch.cuda.Stream(device=device)
        # the exchange runs on its own high-priority stream: its blocks take the SM slots the hash-grid backward's blocks free
        self.x_stream = torch.cuda.Stream(device=device, priority=-1)
        self.exchanged, self.zeroed = torch.cuda.Event(), torch.cuda.Event()
        self.zeroed.record(); self.exchanged.record()   # creates the CUDA event handles
        self.zero_pending = False
        # a wait that times out raises this host-visible word instead of trapping (arn_p2p_set_error_word)
        self.err = torch.zeros(1, dtype=torch.int32).pin_memory()
        _lib.call("arn_p2p_set_error_word", self.err.data_ptr())
        _lib.call("arn_p2p_set_timeout", float(os.environ.get("ARN_P2P_TIMEOUT_S", "120")))
        self.grid_overlap = int(os.environ.get("ARN_P2P_GRID_OVERLAP", "2"))   # blocks per SM while the backward of the next group runs
        self.grid_tail = int(os.environ.get("ARN_P2P_GRID_TAIL", "8"))
        dist.barrier()  # every mapping exists before anyone signals

    def check(self):
        """Raises if a peer wait timed out since the last look (host read of a pinned word: no synchronisation)."""
        code = int(self.err[0])
        if code:
            raise RuntimeError(f"peer-memory exchange: rank {self.rank} waited longer than the timeout for rank {code & 0xff} "
                               f"({'barrier' if code & 0x200 else 'wait'}); a rank died or fell far behind -- put a dist.barrier() in front of "
                               "rank-asymmetric work, or raise ARN_P2P_TIMEOUT_S")

    def _exchange_group(self, g, p_flat, m, v, hyper, cuda_stream):
        from ._lib import call, ptr
        lo, cnt, off = self.groups[g]
        if cnt > 0 and self.mc_grad is not None:
            call("arn_p2p_adam_exchange_mc", self.mc_grad, self.mc_p16, lo, cnt, ptr(p_flat[lo:lo + cnt]), ptr(m[off:off + cnt]), ptr(v[off:off + cnt]),
                 *hyper, cuda_stream)
        elif cnt > 0:
            call("arn_p2p_adam_exchange", self.G, self.H, self.world, lo, cnt, ptr(p_flat[lo:lo + cnt]), ptr(m[off:off + cnt]), ptr(v[off:off + cnt]),
                 *hyper, cuda_stream)

    def step(self, p_flat, m, v, hyper, step_id, cuda_stream, group_events=None):
        """hyper = (lr, beta1, beta2, eps, step, inv_grad_scale).  group_events = None: the gradients of every rank must be
        complete on `cuda_stream`; everything is queued there.  group_events = [torch.cuda.Event per group]: event g is recorded
        (arn_train_set_level_groups) where group g's gradient is final on this rank; the exchange of group g then runs on the
        exchange stream behind a flag barrier of its own -- beside the backward of group g+1 -- and `cuda_stream` only waits for
        the end of the last one."""
        from ._lib import call
        self.check()
        G = len(self.groups)
        main = torch.cuda.current_stream()
        if group_events is None:
            call("arn_p2p_barrier", self.F, self.f_ptr, self.world, self.rank, 0, step_id, cuda_stream)   # my gradients are final ... and everyone's
            call("arn_p2p_set_grid", self.grid_tail)
            for g in range(G):
                self._exchange_group(g, p_flat, m, v, hyper, cuda_stream)
            # my fp16 slices are in every copy and I am done reading ... my copy is complete, nobody reads my gradients any more
            # (measured: folding the two synchronisations into the exchange kernel -- spinning blocks at its start, a system fence
            # per thread at its end -- is slower than the two one-block barrier launches: 0.420 vs 0.389 ms per step at 2 GPUs)
            call("arn_p2p_barrier", self.F, self.f_ptr, self.world, self.rank, G, step_id, cuda_stream)
            self.exchanged.record(main)
        else:
            assert len(group_events) == G
            xs = self.x_stream
            xh = xs.cuda_stream
            for g in range(G):
                xs.wait_event(group_events[g])
                call("arn_p2p_barrier", self.F, self.f_ptr, self.world, self.rank, g, step_id, xh)        # group g is final on every rank
                call("arn_p2p_set_grid", self.grid_overlap if g < G - 1 else self.grid_tail)
                self._exchange_group(g, p_flat, m, v, hyper, xh)
            call("arn_p2p_barrier", self.F, self.f_ptr, self.world, self.rank, G, step_id, xh)
            self.exchanged.record(xs)
            main.wait_event(self.exchanged)  # whatever reads the table next
        # The 45.8 MB memset runs on its own stream, under the next step's forward; whoever writes gradients next waits for
        # `zeroed` (NGPTrainer arms arn_train_set_join in front of the MLP backward; wait_zeroed() for everybody else).
        self.zero_stream.wait_event(self.exchanged)
        with torch.cuda.stream(self.zero_stream):
            self.grad.zero_()
        self.zeroed.record(self.zero_stream)
        self.zero_pending = True

    def wait_zeroed(self):
        """Make the current stream wait for the gradient buffer's asynchronous zeroing (no-op if none is pending)."""
        if self.zero_pending:
            torch.cuda.current_stream().wait_event(self.zeroed)
            self.zero_pending = False

    @torch.no_grad()
    def gather_owned(self, flat):
        """In place: every rank's owned runs of `flat` (n elements) become current everywhere (checkpoints)."""
        tmp = torch.zeros_like(flat)
        for lo, cnt, _ in self.groups:
            tmp[lo:lo + cnt] = flat[lo:lo + cnt]
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM)  # every element has exactly one owner
        flat.copy_(tmp)
