"""world_size-2 gloo test of the ray-sharding + single all-gather host logic (CPU; the render
function is a stand-in because the real one needs a GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_pl_b200.sharded import render_rays_sharded, shard_bounds


def fake_render(rays, scale):
    return {"rgb_fine": rays[:, :3] * scale, "depth_fine": rays[:, 6] + rays[:, 7], "opacity_fine": rays[:, 0] * 0 + 1}


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        rays = torch.rand(n, 8, generator=g)
        out = render_rays_sharded(fake_render, rays, 2.0)
        ref = fake_render(rays, 2.0)
        assert set(out) == set(ref)
        for k in ref:
            assert out[k].shape == ref[k].shape, (k, out[k].shape)
            assert torch.equal(out[k], ref[k]), k
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds():
    assert shard_bounds(10, 2, 0) == (0, 5, 5) and shard_bounds(10, 2, 1) == (5, 10, 5)
    assert shard_bounds(11, 2, 1) == (6, 11, 6)
    assert shard_bounds(3, 8, 7) == (3, 3, 1)
    cover = sorted(sum(([i for i in range(*shard_bounds(1001, 8, r)[:2])] for r in range(8)), []))
    assert cover == list(range(1001))


def test_sharded_render_world2_even_and_ragged():
    for n in (64, 33, 1):
        mp.spawn(_worker, args=(2, _free_port(), n), nprocs=2, join=True)


def test_single_process_passthrough():
    rays = torch.rand(5, 8)
    out = render_rays_sharded(fake_render, rays, 3.0)
    assert torch.equal(out["rgb_fine"], rays[:, :3] * 3.0)
