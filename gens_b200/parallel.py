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

from typing import List, Optional, Sequence, Tuple

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


class SlabExchange:
    """Peer-mapped output buffers for the fused build + exchange (one per process and pyramid shape).

    Two symmetric allocations (torch.distributed._symmetric_memory: every rank's buffer is mapped into every
    other rank's address space over NVLink) hold the final (1,9,D,D,D) tensors of all scales; builds alternate
    between them.  K1 stores each result into the same offset of ALL ranks' buffers, a device-side barrier on
    the stream closes the build.  OWNERSHIP: the tensors returned by build k are views of buffer k % 2 and are
    overwritten by build k + 2 (the reference consumes a scene's volumes before it builds the next one); a rank
    may not run more than one build ahead of its consumers, which stream order gives for free."""

    _cache = {}

    def __init__(self, dims, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.dims = list(dims)
        self.offs = [0]
        for d in self.dims:
            self.offs.append(self.offs[-1] + 9 * d ** 3)
        self.bufs, self.handles = [], []
        for _ in range(2):
            t = symm_mem.empty(self.offs[-1], dtype=torch.float32, device=device)
            self.handles.append(symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD))
            self.bufs.append(t)
        self.turn = 0

    @classmethod
    def get(cls, dims, device, group):
        key = (tuple(dims), str(device), id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(dims, device, group)
        return cls._cache[key]

    def next(self):
        """(local buffer, its handle) of this build; alternates between the two allocations."""
        i = self.turn
        self.turn ^= 1
        return self.bufs[i], self.handles[i]


def fused_sharded_agg_mean_var(volume_module, features, intrs, c2ws, rank: int, world: int, min_vis_view: int = 1,
                               group=None):
    """Slab-sharded build whose stores ARE the exchange: every rank computes its planes of every scale and K1
    writes them into the final tensors of all ranks over NVLink peer mappings (no all-gather, no scatter
    pass); one device-side barrier makes the assembled volumes visible.  Bit-identical to the 1-GPU build.
    See SlabExchange for the ownership of the returned tensors."""
    from .volume import agg_mean_var
    dims = volume_module.volume_dims
    if world > 8:
        raise RuntimeError("fused slab exchange supports up to 8 ranks of one node")
    ex = SlabExchange.get(dims, features[0].device, group)
    buf, hdl = ex.next()
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    slabs = [slab_bounds(d, rank, world) for d in dims]
    peer_outs = []
    for d, off in zip(dims, ex.offs):
        vol = buf[off: off + 8 * d ** 3].view(1, 8, d, d, d)
        msk = buf[off + 8 * d ** 3: off + 9 * d ** 3].view(1, 1, d, d, d)
        peer_outs.append((vol, msk, [p + 4 * off for p in ptrs], [p + 4 * (off + 8 * d ** 3) for p in ptrs], rank))
    vols, masks = agg_mean_var(features, intrs, c2ws, dims, min_vis_view, slabs, volume_module.div_mode,
                               peer_outs=peer_outs)
    hdl.barrier(channel=0)  # every rank's stores into this buffer have landed (stream-ordered on all ranks)
    # the exchange buffers are written through raw pointers (no _version bump): derived-data caches keyed on tensor
    # versions (channels-last copies, the TV value) must not outlive a build into the same storage
    from . import projector
    projector.clear_caches()
    return vols, masks


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
    if torch.is_grad_enabled() and any(f.requires_grad for f in features[:len(dims)]):
        raise RuntimeError("gens_b200: sharded_agg_mean_var is inference-only (the gathered volumes carry no grad_fn); "
                           "run it under torch.no_grad()")
    vols, masks = agg_mean_var(features, intrs, c2ws, dims, min_vis_view, slabs, volume_module.div_mode)
    full_v = [gather_slabs(v, d, world, group) for v, d in zip(vols, dims)]
    full_m = [gather_slabs(m, d, world, group) for m, d in zip(masks, dims)]
    return full_v, full_m


def gather_slab_volumes(slabs: Sequence[torch.Tensor], rank: int, world: int, group=None) -> List[torch.Tensor]:
    """slabs[i] (1,C,D_i/P,D_i,D_i) on each rank -> (1,C,D_i,D_i,D_i) on every rank (in place on CUDA: the slab is
    copied to its planes of the final tensor and one all-gather per channel fills in the rest)."""
    out = []
    for s in slabs:
        d = s.shape[3]
        if world == 1:
            out.append(s)
        elif s.is_cuda and d % world == 0:
            a0, a1 = slab_bounds(d, rank, world)
            full = torch.empty((1, s.shape[1], d, d, d), device=s.device, dtype=s.dtype)
            full[:, :, a0:a1].copy_(s)
            out.append(gather_slabs_inplace(full, d, rank, world, group))
        else:
            out.append(gather_slabs(s, d, world, group))
    return out


@torch.no_grad()
def sharded_build_and_regularise(volume_module, reg_network, features, intrs, c2ws, rank: int, world: int,
                                 min_vis_view: int = 1, group=None, graphed=None):
    """The volume side of GenS.forward / init_volumes (reference models/gens.py:68-70, :143-145) on P GPUs without
    ever assembling the 9-channel volumes: every rank builds its x-slabs (K1), the regulariser runs slab-parallel on
    them (gens_b200/reg_network.py: one halo plane per layer and side, InstanceNorm statistics all-reduced), and only
    the 4-channel results and the masks are gathered -- 384 MB instead of 690 MB per build for the config-2
    pyramid, and 1/P of the regulariser's work per GPU.  Returns (volumes, mask_volumes) as
    `reg_network(volume.agg_mean_var(...)[0])`, `...[1]` would on one GPU, on every rank.  Inference only.
    `graphed` = a reg_network.PeerSlabRegulariser built for this network and pyramid: halo planes, moments and
    results travel through NVLink peer memory and the whole step replays as one CUDA graph (the returned tensors are then
    its static outputs)."""
    from .volume import agg_mean_var
    dims = volume_module.volume_dims
    slabs = [slab_bounds(d, rank, world) for d in dims]
    if graphed is not None:  # K1 writes into the graph's static inputs, one replay does the rest
        agg_mean_var(features, intrs, c2ws, dims, min_vis_view, slabs, volume_module.div_mode,
                     outs=list(zip(graphed.inputs, graphed.masks)))
        return graphed()
    vols, masks = agg_mean_var(features, intrs, c2ws, dims, min_vis_view, slabs, volume_module.div_mode)
    reg = reg_network.forward_slabs(vols, rank, world, group)
    del vols
    return gather_slab_volumes(reg, rank, world, group), gather_slab_volumes(masks, rank, world, group)


def gather_rays(local: torch.Tensor, n_total: int, rank: int, world: int, group=None) -> torch.Tensor:
    """Per-ray outputs (n_local, ...) of contiguous ray shards -> (n_total, ...) on every rank."""
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    return torch.cat(_all_gather_rows(local, sizes, group), dim=0)


def sharded_sdf_grid(sdf_grid_fn, resolution: int, rank: int, world: int, group=None, dst: Optional[int] = 0):
    """The mesh-extraction lattice (reference implicit_surface.py:407-421) sharded by
    x-slabs: rank r evaluates planes shard_range(resolution, r, world) with
    `sdf_grid_fn(x_range) -> (x1 - x0, resolution, resolution)` (ImplicitSurface.sdf_grid bound to its
    volumes) -- points are independent, so there is no collective during compute -- and the slabs are then
    gathered: on rank `dst` only (the rank that runs marching cubes; the others get None) or, with
    dst=None, on every rank.  Bit-identical to the single-GPU lattice: a point's value does not depend on
    which rank evaluates it."""
    x0, x1 = shard_range(resolution, rank, world)
    local = sdf_grid_fn((x0, x1))
    if tuple(local.shape) != (x1 - x0, resolution, resolution):
        raise RuntimeError(f"sdf_grid_fn returned {tuple(local.shape)} for x_range {(x0, x1)}")
    if world == 1:
        return local
    sizes = [shard_range(resolution, r, world)[1] - shard_range(resolution, r, world)[0] for r in range(world)]
    if dst is None:
        return torch.cat(_all_gather_rows(local.contiguous(), sizes, group), dim=0)
    top = max(sizes)  # equal-sized messages for every backend: pad the short shards, trim after the gather
    send = local.contiguous()
    if send.shape[0] < top:
        send = torch.cat([send, send.new_zeros((top - send.shape[0], resolution, resolution))], dim=0)
    outs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, outs, dst=dst, group=group)
    return torch.cat([o[:n] for o, n in zip(outs, sizes)], dim=0) if rank == dst else None
