"""world_size-2 gloo test of the N>1 host logic (ray sharding, gradient sum, frame gather) on CPU tensors."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ar_nerf_b200.sharding import (all_gather_shards, allreduce_grads, gather_frame, gather_frame_interleaved, group_runs, padded_numel,
                                   reduce_scatter_sum, shard_bounds, shard_rays, shard_rays_interleaved, shard_size)


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 8192, 640000, 640001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_group_runs_tile_every_level_group():
    """Ownership of the level-grouped peer exchange (sharding.group_runs): for every group the ranks' runs are disjoint,
    4-element aligned and cover the group exactly; the packed optimizer state of a rank is the concatenation of its runs."""
    from ar_nerf_b200.field import HashGeometry
    geo = HashGeometry()
    n = 3072 + 2 * geo.total
    for levels in ([0, 16], [0, 8, 16], [0, 8, 11, 13, 16], list(range(17))):
        bounds = [0] + [3072 + 2 * int(geo.offset[l]) for l in levels[1:-1]] + [n]
        for world in (1, 2, 4, 8):
            owner = torch.full((n,), -1, dtype=torch.int8)
            for rank in range(world):
                off_expect = 0
                for (lo, cnt, off), a, b in zip(group_runs(bounds, world, rank), bounds, bounds[1:]):
                    assert lo % 4 == 0 and cnt % 4 == 0 and a <= lo and lo + cnt <= b and off == off_expect
                    assert (owner[lo:lo + cnt] == -1).all()
                    owner[lo:lo + cnt] = rank
                    off_expect += cnt
            assert (owner >= 0).all()


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    rays_o = torch.rand(1001, 3, generator=g); rays_d = torch.rand(1001, 3, generator=g)
    o, d = shard_rays(rays_o, rays_d, rank, world)
    lo, hi = shard_bounds(1001, rank, world)
    assert torch.equal(o, rays_o[lo:hi]) and torch.equal(d, rays_d[lo:hi])
    # per-rank "gradient" = sum of its rays; the all-reduced sum equals the single-process gradient
    grad = [o.sum(0).clone(), d.sum(0).clone()]
    allreduce_grads(grad, world)
    assert torch.allclose(grad[0], rays_o.sum(0), atol=1e-4) and torch.allclose(grad[1], rays_d.sum(0), atol=1e-4)
    # sharded optimizer plumbing: reduce-scatter of a padded gradient, a slice-local update, all-gather of the slices
    n = 1003
    S, P = shard_size(n, world), padded_numel(n, world)
    assert S % 8 == 0 and P == S * world and P >= n
    gfull = torch.zeros(P); gfull[:n] = torch.arange(n, dtype=torch.float32) * (rank + 1)
    gs = torch.empty(S)
    reduce_scatter_sum(gfull, gs, rank, world)
    want = torch.zeros(P); want[:n] = torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world))
    assert torch.equal(gs, want[rank * S:(rank + 1) * S])
    params = torch.zeros(P); params[rank * S:(rank + 1) * S] = -0.5 * gs   # each rank updates its slice only
    all_gather_shards(params, rank, world)
    assert torch.equal(params, -0.5 * want)
    oi, di = shard_rays_interleaved(rays_o, rays_d, rank, world)
    assert torch.equal(oi, rays_o[rank::world]) and torch.equal(di, rays_d[rank::world])
    fi = gather_frame_interleaved(oi * 3, 1001, rank, world)
    assert (fi is None) if rank else torch.equal(fi, rays_o * 3)
    frame = gather_frame(o * 2, 1001, rank, world)
    if rank == 0:
        assert torch.equal(frame, rays_o * 2)
    else:
        assert frame is None
    dist.destroy_process_group()


def test_two_rank_gloo():
    mp.spawn(_worker, args=(2, 29533), nprocs=2, join=True)
