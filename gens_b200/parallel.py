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


def gather_slabs_inplace(full: torch.Tensor, d: int, rank: int, world: int, group=None) -> torch.Tensor:
    """`full` (1,C,D,D,D) already holds this rank's planes [a0,a1) of every channel; fill in the rest.
    In NCDHW every channel plane is the rank-ordered concatenation of the slabs, so one IN-PLACE
    all-gather per channel (input = the rank's chunk of the output) assembles the tensor with no staging
    copy; the C x S collectives of a build are issued as one coalesced NCCL group."""
    if world == 1:
        return full
    if d % world != 0:
        raise RuntimeError("in-place slab gather needs D divisible by the number of ranks")
    a0, a1 = slab_bounds(d, rank, world)
    for c in range(full.shape[1]):
        dist.all_gather_into_tensor(full[0, c], full[0, c, a0:a1], group=group)
    return full


def sharded_agg_mean_var(volume_module, features, intrs, c2ws, rank: int, world: int, min_vis_view: int = 1,
                         group=None):
    """Slab-sharded Volume.agg_mean_var + all-gather; same return value as the single-GPU call.

    When every D divides by the world size, K1 writes all scales' slabs (8 channels + mask each) straight
    into ONE contiguous send buffer, a single all-gather moves it, and one scatter kernel per scale
    (gens_unpack_slabs) lays the rank-major blocks out as the final NCDHW tensors."""
    from . import _lib
    from .volume import agg_mean_var
    dims = volume_module.volume_dims
    slabs = [slab_bounds(d, rank, world) for d in dims]
    if all(d % world == 0 and (d * d) % 4 == 0 for d in dims) and features[0].is_cuda:
        dev = features[0].device
        sizes = [9 * (d // world) * d * d for d in dims]
        offs = [0]
        for sz in sizes:
            offs.append(offs[-1] + (sz + 3) // 4 * 4)
        stride = offs[-1]
        send = torch.empty(stride, device=dev, dtype=torch.float32)
        outs = []
        for d, off in zip(dims, offs):
            p = d // world
            outs.append((send[off: off + 8 * p * d * d].view(1, 8, p, d, d),
                         send[off + 8 * p * d * d: off + 9 * p * d * d].view(1, 1, p, d, d)))
        agg_mean_var(features, intrs, c2ws, dims, min_vis_view, slabs, volume_module.div_mode, outs=outs)
        recv = torch.empty(world * stride, device=dev, dtype=torch.float32)
        dist.all_gather_into_tensor(recv, send, group=group)
        vols = [torch.empty((1, 8, d, d, d), device=dev, dtype=torch.float32) for d in dims]
        masks = [torch.empty((1, 1, d, d, d), device=dev, dtype=torch.float32) for d in dims]
        for d, off, v, m in zip(dims, offs, vols, masks):
            _lib.check(_lib.lib().gens_unpack_slabs(_lib.ptr(recv), world, stride, off, d, _lib.ptr(v), _lib.ptr(m),
                                                    _lib.stream_ptr(dev)), "gens_unpack_slabs")
        return vols, masks
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
