"""Host-side multi-GPU logic: rays are independent, so training shards a global batch (or draws
per-rank batches) with no data-path collective except ONE gradient exchange per step; test rendering shards the frame
into contiguous row bands and gathers the per-ray results on rank 0."""
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
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)
