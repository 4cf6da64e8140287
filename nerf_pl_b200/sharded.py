"""Ray-sharded rendering across the GPUs of one node.

Rays are independent units (no cross-ray term anywhere in reference models/rendering.py), so a
ray batch is split into contiguous equal shards, one per rank (one process per GPU, launched by
torchrun), each rank renders its shard with the fused kernel and the rendered pixels are
exchanged with ONE all-gather per result key set (NCCL over NVLink/NVSwitch; `gloo` in the CPU
tests).  Weights are replicated (2.4 MB fp16 image per network).  The reference has no
equivalent: it renders each image on one GPU (eval.py:117-123).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Contiguous shard [lo, hi) of rank `rank` and the padded per-rank count (ceil(n/world))."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    hi = min(lo + per, n)
    return lo, hi, per


def render_rays_sharded(render_fn: Callable[..., Dict[str, torch.Tensor]], rays: torch.Tensor, *args,
                        group: Optional[dist.ProcessGroup] = None, **kwargs) -> Dict[str, torch.Tensor]:
    """Every rank passes the SAME full `rays` (N, 8); returns the full-size result dict on every
    rank.  `render_fn(rays_shard, *args, **kwargs)` is e.g. a partial of nerf_pl_b200.render_rays.
    All result tensors are packed column-wise into one (per, C) buffer so the exchange is a single
    all_gather regardless of how many keys the result has."""
    if not (dist.is_available() and dist.is_initialized()):
        return render_fn(rays, *args, **kwargs)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = rays.shape[0]
    lo, hi, per = shard_bounds(n, world, rank)
    shard = rays[lo:hi]
    if hi - lo < per:                       # pad the tail shard with copies of a valid ray
        filler = rays[max(n - 1, 0):n].expand(per - (hi - lo), rays.shape[1])
        shard = torch.cat((shard, filler), 0)
    out = render_fn(shard.contiguous(), *args, **kwargs)
    keys = sorted(out)
    cols = [out[k].reshape(per, -1).to(torch.float32) for k in keys]
    widths = [c.shape[1] for c in cols]
    packed = torch.cat(cols, 1).contiguous()
    gathered = torch.empty(world * per, packed.shape[1], dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    gathered = gathered[:n]
    result, c0 = {}, 0
    for k, wdt in zip(keys, widths):
        result[k] = gathered[:, c0:c0 + wdt].reshape((n,) + tuple(out[k].shape[1:]))
        c0 += wdt
    return result
