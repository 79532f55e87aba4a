"""Single-node multi-GPU sharding of the hot path (one process per GPU, torch.distributed).

  * volume build: every rank builds the planes [a0,a1) of tensor dim 2 (world x) of every scale with
    the same K1 launch it would use alone (`slabs=`), then ONE all-gather per scale assembles the full
    (1,8,D,D,D) / (1,1,D,D,D) tensors on every rank (NCCL over NVLink on GPUs, gloo in the CPU tests).
    The gathered result is bit-identical to the single-GPU build: slabs are computed independently.
  * ray marching: rays are split into contiguous ranges, no collective during compute, one gather of
    the per-ray outputs at the end.
The reference has neither (it is data-parallel over scenes only, runner.py:104).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def slab_bounds(d: int, rank: int, world: int) -> Tuple[int, int]:
    """Planes [a0,a1) of a D-plane axis owned by `rank` (contiguous, sizes differ by at most one)."""
    return (d * rank) // world, (d * (rank + 1)) // world


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    return (n * rank) // world, (n * (rank + 1)) // world


def _all_gather_rows(local: torch.Tensor, sizes: Sequence[int], group=None) -> List[torch.Tensor]:
    """All-gather along dim 0 with per-rank row counts `sizes` (padded to the largest shard so that every
    backend -- NCCL, gloo -- sees equal-sized messages).  Returns the per-rank pieces, trimmed."""
    world, top = len(sizes), max(sizes)
    send = local.contiguous()
    if send.shape[0] < top:
        pad = torch.zeros((top - send.shape[0],) + tuple(send.shape[1:]), device=send.device, dtype=send.dtype)
        send = torch.cat([send, pad], dim=0)
    buf = torch.empty((world * top,) + tuple(send.shape[1:]), device=send.device, dtype=send.dtype)
    dist.all_gather_into_tensor(buf, send, group=group)
    return [buf[r * top: r * top + sizes[r]] for r in range(world)]


def gather_slabs(slab: torch.Tensor, d: int, world: int, group=None) -> torch.Tensor:
    """slab (1,C,planes_r,D,D) on each rank -> (1,C,D,D,D) on every rank."""
    if world == 1:
        return slab
    c = slab.shape[1]
    sizes = [slab_bounds(d, r, world)[1] - slab_bounds(d, r, world)[0] for r in range(world)]
    # plane-major messages; then one strided copy into the final channel-major NCDHW layout
    pieces = _all_gather_rows(slab[0].permute(1, 0, 2, 3), sizes, group)  # each (planes_r, C, D, D)
    return torch.cat(pieces, dim=0).permute(1, 0, 2, 3).reshape(1, c, d, d, d).contiguous()


def sharded_agg_mean_var(volume_module, features, intrs, c2ws, rank: int, world: int, min_vis_view: int = 1,
                         group=None):
    """Slab-sharded Volume.agg_mean_var + all-gather; same return value as the single-GPU call."""
    from .volume import agg_mean_var
    dims = volume_module.volume_dims
    slabs = [slab_bounds(d, rank, world) for d in dims]
    vols, masks = agg_mean_var(features, intrs, c2ws, dims, min_vis_view, slabs, volume_module.div_mode)
    full_v = [gather_slabs(v, d, world, group) for v, d in zip(vols, dims)]
    full_m = [gather_slabs(m, d, world, group) for m, d in zip(masks, dims)]
    return full_v, full_m


def gather_rays(local: torch.Tensor, n_total: int, rank: int, world: int, group=None) -> torch.Tensor:
    """Per-ray outputs (n_local, ...) of contiguous ray shards -> (n_total, ...) on every rank."""
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    return torch.cat(_all_gather_rows(local, sizes, group), dim=0)
