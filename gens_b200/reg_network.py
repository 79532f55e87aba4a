"""The volume regulariser that consumes K1's output, and its slab-parallel form.

Reference: models/modules/reg_network.py:105-166 (`RegNetwork`): a 3-D U-Net over the five (1,8,D_i,D_i,D_i) mean/var
volumes -- conv0, five stride-2 encoder stages (the next coarser volume is concatenated after each), five transposed-
convolution decoder stages with skip additions, one 3x3x3 output convolution per scale -> five (1,4,D_i,D_i,D_i)
volumes that the ray marcher samples.  Every convolution except the output ones is followed by InstanceNorm3d
(no affine, eps 1e-5) and ReLU.  `RegNetwork` here keeps the reference's parameter names, so its state_dict loads.

Why it is in this repo: on P GPUs the reference's data flow needs the full 9-channel volumes on every rank BEFORE this
network (690 MB ingested per GPU, the floor of the fused slab exchange of gens_b200/parallel.py).  `forward_slabs` runs
the SAME network on the x-slabs K1 leaves on each rank instead:
  * a 3x3x3 convolution needs one neighbouring plane per side (stride 2: only the lower side; transposed stride 2: only
    the upper side) -- one plane exchanged with each neighbour per layer (`_Ops.halo`, point-to-point);
  * InstanceNorm needs the per-channel mean / variance of the WHOLE volume -- every rank reduces its slab and one
    all-reduce of 2C numbers per layer combines them (parallel-variance formula, accumulated in float64);
  * only the 4-channel results (and the 1-channel masks) are gathered afterwards: 384 MB instead of 690 MB per
    build, and the network itself runs on 1/P of the voxels per GPU.
Two transports for those exchanges: NCCL messages (`_Ops`: point-to-point halo planes, a 2C-double all-reduce per
layer, all-gathers of the results; also what the gloo tests run) and NVLink peer memory (`_PeerOps` /
`PeerSlabRegulariser`: layer outputs in a symmetric arena, halos and moments read from the neighbours in place, device
barriers, the whole step replayed as one CUDA graph per rank).
Kernels: the layers with up to 16 output channels (stride 1, stride 2, transposed stride 2) and every InstanceNorm run
on K13 (csrc/conv3d.cu) -- in cuDNN / ATen those take ~105 of the network's 112 ms on one B200; the deep narrow layers
(32^3 and coarser) stay cuDNN (library calls, counted as such).  The slab path is inference only (like the other sharded
paths): the reference trains with one scene per GPU (DDP replicas), and with autograd enabled `forward` is the
reference's own op sequence.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import ctypes

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

EPS = 1e-5  # nn.InstanceNorm3d default, reg_network.py:16


class _Unit(nn.Module):
    """Holder of one normalised convolution; the weight lives in `.conv` as in the reference's Conv3d / Deconv3d
    blocks (reg_network.py:7-50), bias-free because a normalisation follows."""

    def __init__(self, c_in: int, c_out: int, stride: int = 1, transposed: bool = False):
        super().__init__()
        self.stride, self.transposed = stride, transposed
        if transposed:
            self.conv = nn.ConvTranspose3d(c_in, c_out, 3, stride=stride, padding=1, output_padding=stride - 1, bias=False)
        else:
            self.conv = nn.Conv3d(c_in, c_out, 3, stride=stride, padding=1, bias=False)


class _LocalOps:
    """Whole volumes on one device."""

    @staticmethod
    def unit(x: torch.Tensor, u: _Unit, skip: Optional[torch.Tensor] = None) -> torch.Tensor:
        if u.transposed:
            y = F.conv_transpose3d(x, u.conv.weight, None, stride=u.stride, padding=1, output_padding=u.stride - 1)
        else:
            y = F.conv3d(x, u.conv.weight, None, stride=u.stride, padding=1)
        y = F.relu_(F.instance_norm(y, eps=EPS))
        return y if skip is None else y + skip

    @staticmethod
    def out(x: torch.Tensor, conv: nn.Conv3d) -> torch.Tensor:
        return F.conv3d(x, conv.weight, conv.bias, stride=1, padding=1)

    @staticmethod
    def cat(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return torch.cat([a, b], dim=1)


class _Ops:
    """Inference ops on x-slabs (planes of tensor dim 2) of every tensor on `world` ranks; rank r owns planes
    [r d/P, (r+1) d/P).  world = 1: whole volumes.  On CUDA the layers with c_in % 8 == 0 and up to 16 output channels
    run on K13 (csrc/conv3d.cu: direct FFMA2 convolutions -- stride 1, stride 2, transposed stride 2 -- that also reduce
    the InstanceNorm moments and read the halo planes through their own pointers) and every normalisation on its
    in-place kernel; the deep narrow levels are cuDNN with the moments taken by torch.var_mean."""

    def __init__(self, rank: int = 0, world: int = 1, group=None):
        self.rank, self.world, self.group = rank, world, group
        self._packed = {}

    def _peer(self, r: int) -> int:
        return r if self.group is None else dist.get_global_rank(self.group, r)

    # -- hooks the peer-memory variant overrides ----------------------------------------------------------------
    def _new(self, shape, device) -> torch.Tensor:
        """Storage of a layer's output."""
        return torch.empty(shape, device=device, dtype=torch.float32)

    def _own(self, y: torch.Tensor) -> torch.Tensor:
        """A library-produced tensor as a layer output (the peer variant copies it to where neighbours can read it)."""
        return y

    def _combine(self, stats: torch.Tensor) -> torch.Tensor:
        """The volume's moments from this rank's."""
        if self.world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
        return stats

    def _publish(self) -> None:
        """Called once a layer's output is final."""

    def cat(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return torch.cat([a, b], dim=1)

    def halo(self, x: torch.Tensor, lower: bool, upper: bool):
        """The neighbours' boundary planes: (plane below the slab or None, plane above or None); None at the
        volume's ends (= the convolution's zero padding)."""
        r, p = self.rank, self.world
        lo = torch.empty_like(x[:, :, :1]) if lower and r > 0 else None
        hi = torch.empty_like(x[:, :, :1]) if upper and r + 1 < p else None
        ops, keep = [], []
        if lower and r + 1 < p:                     # my last plane is the lower halo of rank r+1
            keep.append(x[:, :, -1:].contiguous())
            ops.append(dist.P2POp(dist.isend, keep[-1], self._peer(r + 1), self.group))
        if lo is not None:
            ops.append(dist.P2POp(dist.irecv, lo, self._peer(r - 1), self.group))
        if upper and r > 0:                         # my first plane is the upper halo of rank r-1
            keep.append(x[:, :, :1].contiguous())
            ops.append(dist.P2POp(dist.isend, keep[-1], self._peer(r - 1), self.group))
        if hi is not None:
            ops.append(dist.P2POp(dist.irecv, hi, self._peer(r + 1), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return lo, hi

    @staticmethod
    def _padded(x, lo, hi, lower: bool, upper: bool):
        parts = ([lo if lo is not None else torch.zeros_like(x[:, :, :1])] if lower else []) + [x] + \
                ([hi if hi is not None else torch.zeros_like(x[:, :, :1])] if upper else [])
        return torch.cat(parts, 2) if len(parts) > 1 else x

    def _k13(self, x, conv) -> bool:
        w = conv.weight
        return (x.is_cuda and x.dtype == torch.float32 and isinstance(conv, nn.Conv3d) and conv.stride == (1, 1, 1)
                and w.shape[1] % 8 == 0 and w.shape[0] in (4, 8, 16) and x.shape[0] == 1)

    def _conv_k13(self, x, conv, lo, hi, want_stats: bool):
        from . import _lib
        w = conv.weight
        key = (w.data_ptr(), w._version)
        pk = self._packed.get(key)
        if pk is None:  # (c_out, c_in, kd, kh, kw) -> [c_in][kh][kw][kd][c_out]
            pk = self._packed[key] = w.detach().permute(1, 3, 4, 2, 0).contiguous().float()
        x = _lib.f32c(x)
        _, c_in, d, h, wd = x.shape
        c_out = w.shape[0]
        y = self._new((1, c_out, d, h, wd), x.device)
        stats = torch.zeros(2 * c_out, device=x.device, dtype=torch.float64) if want_stats else None
        null = ctypes.c_void_p(0)
        # contiguous copies of the halo planes (peer-memory views are strided) must BOTH stay referenced until the launch:
        # a temporary released between two argument evaluations hands its block to the next one
        lo_c = lo.contiguous() if lo is not None else None
        hi_c = hi.contiguous() if hi is not None else None
        _lib.check(_lib.lib().gens_conv3d_k3(
            _lib.ptr(x), _lib.ptr(lo_c) if lo_c is not None else null,
            _lib.ptr(hi_c) if hi_c is not None else null, _lib.ptr(pk),
            _lib.ptr(conv.bias.detach()) if conv.bias is not None else null, c_in, c_out, d, h, wd, _lib.ptr(y),
            _lib.ptr(stats) if want_stats else null, _lib.stream_ptr(x.device)), "gens_conv3d_k3")
        return y, stats

    def norm_relu_(self, y: torch.Tensor, stats: Optional[torch.Tensor], skip: Optional[torch.Tensor] = None):
        """InstanceNorm over the WHOLE volume + ReLU (+ skip), in place.  stats = per-channel [sum | sum of squares] of
        this rank's slab (float64); None: taken here.  Slabs are equal-sized, so the all-reduced sums over P x the local
        count are the volume's moments."""
        c = y.shape[1]
        n_local = y[0, 0].numel()
        if stats is None:
            var, mean = torch.var_mean(y, dim=(0, 2, 3, 4), unbiased=False)
            mean, var = mean.double(), var.double()
            stats = torch.cat([mean, var + mean * mean]) * n_local
        stats = self._combine(stats)
        count = float(n_local * self.world)
        if y.is_cuda and y.dtype == torch.float32 and y.is_contiguous() and n_local % 4 == 0 and \
                (skip is None or (skip.is_contiguous() and skip.dtype == torch.float32)):
            from . import _lib
            _lib.check(_lib.lib().gens_instnorm_relu(
                _lib.ptr(y), _lib.ptr(stats), c, n_local, count, EPS,
                _lib.ptr(skip) if skip is not None else ctypes.c_void_p(0), _lib.stream_ptr(y.device)), "gens_instnorm_relu")
            self._publish()
            return y
        g_mean = stats[:c] / count
        rstd = torch.rsqrt((stats[c:] / count - g_mean * g_mean).clamp_min_(0.0) + EPS)
        shape = (1, -1, 1, 1, 1)
        y = y.sub_(g_mean.float().view(shape)).mul_(rstd.float().view(shape)).relu_()
        y = y.add_(skip) if skip is not None else y
        self._publish()
        return y

    def _strided_k13(self, x, u: _Unit) -> bool:
        w = u.conv.weight
        if not (x.is_cuda and x.dtype == torch.float32 and x.shape[0] == 1):
            return False
        if u.transposed:   # weight (c_in, c_out, 3, 3, 3)
            return w.shape[1] == 8 and w.shape[0] * 27 * 8 * 4 <= 96 * 1024
        return w.shape[0] in (8, 16) and w.shape[1] * 27 * w.shape[0] * 4 <= 96 * 1024 and \
            all(n % 2 == 0 for n in x.shape[2:])

    def _conv_strided_k13(self, x, u: _Unit, halo):
        from . import _lib
        w = u.conv.weight
        key = (w.data_ptr(), w._version)
        pk = self._packed.get(key)
        if pk is None:  # transposed: (c_in, c_out, kd, kh, kw) -> [c_in][kd][kh][kw][c_out]; else as the stride-1 layers
            pk = w.detach().permute(0, 2, 3, 4, 1) if u.transposed else w.detach().permute(1, 3, 4, 2, 0)
            pk = self._packed[key] = pk.contiguous().float()
        x = _lib.f32c(x)
        _, c_in, d, h, wd = x.shape
        c_out = w.shape[1] if u.transposed else w.shape[0]
        shape = (2 * d, 2 * h, 2 * wd) if u.transposed else (d // 2, h // 2, wd // 2)
        y = self._new((1, c_out) + shape, x.device)
        stats = torch.zeros(2 * c_out, device=x.device, dtype=torch.float64)
        fn = _lib.lib().gens_deconv3d_k3s2 if u.transposed else _lib.lib().gens_conv3d_k3s2
        halo_c = halo.contiguous() if halo is not None else None  # referenced until the launch
        _lib.check(fn(_lib.ptr(x), _lib.ptr(halo_c) if halo_c is not None else ctypes.c_void_p(0), _lib.ptr(pk),
                      c_in, c_out, d, h, wd, _lib.ptr(y), _lib.ptr(stats), _lib.stream_ptr(x.device)),
                   "gens_deconv3d_k3s2" if u.transposed else "gens_conv3d_k3s2")
        return y, stats

    def unit(self, x: torch.Tensor, u: _Unit, skip: Optional[torch.Tensor] = None) -> torch.Tensor:
        d = x.shape[2]
        w = u.conv.weight
        stats = None
        if u.transposed:                            # out plane 2i <- in i; out 2i+1 <- in i and i+1: upper halo only
            lo, hi = self.halo(x, False, True) if self.world > 1 else (None, None)
            if self._strided_k13(x, u):
                y, stats = self._conv_strided_k13(x, u, hi)
            else:
                y = self._own(F.conv_transpose3d(self._padded(x, None, hi, False, self.world > 1), w, None, stride=2,
                                                 padding=1, output_padding=1)[:, :, : 2 * d].contiguous())
        elif u.stride == 2:                         # out plane o <- in 2o-1, 2o, 2o+1: lower halo only
            if d % 2:
                raise RuntimeError("slab-parallel RegNetwork: a stride-2 stage met a slab with an odd plane count")
            lo, hi = self.halo(x, True, False) if self.world > 1 else (None, None)
            if self._strided_k13(x, u):
                y, stats = self._conv_strided_k13(x, u, lo)
            else:
                y = self._own(F.conv3d(self._padded(x, lo, None, True, False), w, None, stride=2, padding=(0, 1, 1)))
        else:
            lo, hi = self.halo(x, True, True) if self.world > 1 else (None, None)
            if self._k13(x, u.conv):
                y, stats = self._conv_k13(x, u.conv, lo, hi, True)
            else:
                y = self._own(F.conv3d(self._padded(x, lo, hi, True, True), w, None, stride=1, padding=(0, 1, 1)))
        return self.norm_relu_(y, stats, skip)

    def out(self, x: torch.Tensor, conv: nn.Conv3d) -> torch.Tensor:
        lo, hi = self.halo(x, True, True) if self.world > 1 else (None, None)
        if self._k13(x, conv):
            return self._conv_k13(x, conv, lo, hi, False)[0]
        return F.conv3d(self._padded(x, lo, hi, True, True), conv.weight, conv.bias, stride=1, padding=(0, 1, 1))


class RegNetwork(nn.Module):
    """Same constructor, parameter names and forward as the reference's RegNetwork (reg_network.py:105-166)."""

    def __init__(self, conf=None, d_voluem: Optional[Sequence[int]] = None, d_base: int = 8,
                 d_out: Optional[Sequence[int]] = None):
        super().__init__()
        d_voluem = list(conf.get_list("d_voluem")) if conf is not None else list(d_voluem or [8] * 5)
        d_base = conf.get_int("d_base") if conf is not None else d_base
        d_out = list(conf.get_list("d_out")) if conf is not None else list(d_out or [4] * 5)
        self.num_stage = n = len(d_out)
        width = [d_base * 2 ** i for i in range(n)]          # encoder stage i
        below = [d_base * 2 ** max(i - 1, 0) for i in range(n)]  # what decoder stage i returns to
        self.conv0 = _Unit(d_voluem[0], d_base)
        enc, dec, outs = [], [], []
        c_in = d_base
        for i in range(n):
            enc.append(nn.Sequential(_Unit(c_in, width[i], stride=2), _Unit(width[i], width[i])))
            if i + 1 < n:
                c_in = width[i] + d_voluem[i + 1]
            outs.append(nn.Conv3d(below[i], d_out[i], 3, 1, 1))
            dec.append(_Unit(width[i], below[i], stride=2, transposed=True))
        self.encoder_layers, self.decoder_layers, self.out_layers = nn.ModuleList(enc), nn.ModuleList(dec), nn.ModuleList(outs)

    def _run(self, volumes: Sequence[torch.Tensor], ops) -> List[torch.Tensor]:
        n = self.num_stage
        if len(volumes) != n:
            raise ValueError(f"RegNetwork: expected {n} volumes, got {len(volumes)}")
        e = ops.unit(volumes[0], self.conv0)
        skips = [e]
        for i, stage in enumerate(self.encoder_layers):
            e = ops.unit(ops.unit(e, stage[0]), stage[1])
            skips.append(e)
            if i + 1 < n:
                e = ops.cat(e, volumes[i + 1])
        fine = [None] * n
        d = e
        for i in range(n - 1, -1, -1):
            d = ops.unit(d, self.decoder_layers[i], skips[i])
            fine[i] = d
        return [ops.out(fine[i], self.out_layers[i]) for i in range(n)]

    def forward(self, volumes: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """As the reference's forward.  With autograd enabled (training) or on host tensors: the reference's own op
        sequence (cuDNN / ATen), bit-identical to it on the CPU; under no_grad on CUDA: the K13 path of `_Ops`."""
        if volumes[0].is_cuda and not torch.is_grad_enabled():
            return self._run(volumes, _Ops())
        return self._run(volumes, _LocalOps)

    @torch.no_grad()
    def forward_slabs(self, slabs: Sequence[torch.Tensor], rank: int, world: int, group=None) -> List[torch.Tensor]:
        """`slabs[i]` = planes [r D_i/P, (r+1) D_i/P) of volume i (tensor dim 2) -> the same planes of every output.
        Needs D_i / 2 divisible by P at every scale (so that each stride-2 stage maps slabs onto slabs)."""
        if world == 1:
            return self._run(slabs, _Ops())
        for v in slabs:  # every input scale feeds a stride-2 stage: its slab must hold an even number of planes
            planes, full = v.shape[2], v.shape[3]
            if planes * world != full or planes % 2:
                raise RuntimeError(f"slab-parallel RegNetwork: a {full}^3 volume cannot be cut into {world} slabs "
                                   "with an even number of planes each")
        return self._run(slabs, _Ops(rank, world, group))


class _Arena:
    """Bump allocator over ONE symmetric allocation (torch.distributed._symmetric_memory): every rank allocates the
    same sequence, so a tensor lives at the same offset of every rank's buffer and `peer(r, t)` is rank r's copy of
    `t`, readable over NVLink like local memory."""

    def __init__(self, n_bytes: int, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.buf = symm_mem.empty(n_bytes // 4, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.base, self.used, self.size = self.buf.data_ptr(), 0, n_bytes

    def alloc(self, shape, dtype=torch.float32) -> torch.Tensor:
        n = 1
        for v in shape:
            n *= int(v)
        item = torch.empty((), dtype=dtype).element_size()
        off = (self.used + 255) // 256 * 256
        if off + n * item > self.size:
            raise RuntimeError("slab regulariser: symmetric arena exhausted")
        self.used = off + n * item
        return self.hdl.get_buffer(self.hdl.rank, tuple(int(v) for v in shape), dtype, off // item)

    def peer(self, r: int, t: torch.Tensor) -> torch.Tensor:
        """Rank r's copy of the (contiguous, arena-allocated) tensor or leading slice `t`."""
        off = t.data_ptr() - self.base
        assert 0 <= off < self.size and t.is_contiguous()
        return self.hdl.get_buffer(r, tuple(t.shape), t.dtype, off // t.element_size())

    def barrier(self):
        self.hdl.barrier(channel=1)


class _PeerOps(_Ops):
    """The slab ops with every exchange done through peer memory instead of NCCL: layer outputs live in a symmetric
    arena, a halo plane is the neighbour's tensor read in place over NVLink by the consuming kernel, the InstanceNorm
    moments of all ranks are read from their arenas and summed locally, and one device-side barrier per step orders
    producers and consumers.  No message, no copy, no host synchronisation: the whole forward is graph-capturable."""

    def __init__(self, arena: _Arena, rank: int, world: int):
        super().__init__(rank, world, None)
        self.arena = arena

    def _new(self, shape, device):
        return self.arena.alloc(shape)

    def _own(self, y):
        out = self.arena.alloc(y.shape)
        out.copy_(y)
        return out

    def cat(self, a, b):
        out = self.arena.alloc((1, a.shape[1] + b.shape[1]) + tuple(a.shape[2:]))
        torch.cat([a, b], dim=1, out=out)
        self.arena.barrier()  # the concatenation is a layer input the neighbours read halos of
        return out

    def halo(self, x, lower: bool, upper: bool):
        # x is final on every rank (barrier after the step that produced it); x[0, :, k] is not contiguous across
        # channels, so the neighbour's WHOLE tensor is viewed and the plane sliced out of it
        r, p = self.rank, self.world
        lo = self.arena.peer(r - 1, x)[:, :, -1:] if lower and r > 0 else None
        hi = self.arena.peer(r + 1, x)[:, :, :1] if upper and r + 1 < p else None
        return lo, hi

    def _combine(self, stats):
        mine = self.arena.alloc(stats.shape, torch.float64)
        mine.copy_(stats)
        self.arena.barrier()  # every rank's moments are written (and everyone is done reading the input's halos)
        return torch.stack([self.arena.peer(r, mine) for r in range(self.world)]).sum(0)  # two launches, not P

    def _publish(self):
        self.arena.barrier()


class PeerSlabRegulariser:
    """K1's slabs -> slab-parallel RegNetwork -> full 4-channel volumes + masks on every rank, all exchanges through
    NVLink peer memory (see _PeerOps), replayed as ONE CUDA graph per rank when capture succeeds.

    Static buffers: `inputs[i]` / `masks[i]` are where K1 must write this rank's slabs (pass them as `outs=` to
    agg_mean_var); `volumes` / `mask_volumes` are the assembled results, overwritten by the next call.  Every rank must
    construct and call it collectively; a rank may not start call k + 1 before all ranks finished call k (the closing
    barrier of a call gives that in stream order)."""

    def __init__(self, net: "RegNetwork", dims: Sequence[int], rank: int, world: int, device, group=None,
                 c_in: int = 8, use_graph: bool = True):
        from . import parallel
        self.net, self.rank, self.world, self.dims = net, rank, world, list(dims)
        per_rank = sum(d ** 3 for d in dims) // world
        self.arena = _Arena(int(per_rank * 4 * 64) + (64 << 20), device, group)   # ~60 channel-volumes of activations
        self.inputs = [self.arena.alloc((1, c_in, d // world, d, d)) for d in dims]
        self.masks = [self.arena.alloc((1, 1, d // world, d, d)) for d in dims]
        self._mark = self.arena.used
        c_out = [net.out_layers[i].weight.shape[0] for i in range(len(dims))]
        self.volumes = [torch.empty((1, c, d, d, d), device=device) for c, d in zip(c_out, dims)]
        self.mask_volumes = [torch.empty((1, 1, d, d, d), device=device) for d in dims]
        self.graph = None
        with torch.no_grad():
            self._run()  # warm-up: cuDNN plans, packed weights
            torch.cuda.synchronize(device)
            if use_graph:
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._run()
                    self.graph = g
                except Exception as exc:  # noqa: BLE001 -- eager peer path keeps working
                    import sys
                    sys.stderr.write(f"gens_b200: CUDA-graph capture of the slab regulariser failed ({exc}); running eagerly\n")
                    torch.cuda.synchronize(device)

    def _run(self):
        from .parallel import slab_bounds
        self.arena.used = self._mark
        ops = _PeerOps(self.arena, self.rank, self.world)
        ops._packed = getattr(self, "_packed", {})
        self._packed = ops._packed
        self.arena.barrier()  # K1 of every rank has written its inputs
        outs = self.net._run(self.inputs, ops)
        outs = [o if o.data_ptr() >= self.arena.base and o.data_ptr() < self.arena.base + self.arena.size else ops._own(o)
                for o in outs]
        self.arena.barrier()  # every rank's results are final
        for slabs, full in ((outs, self.volumes), (self.masks, self.mask_volumes)):
            for s, f, d in zip(slabs, full, self.dims):
                for r in range(self.world):  # pull every rank's slab (own included) into the full tensor
                    a0, a1 = slab_bounds(d, r, self.world)
                    f[:, :, a0:a1].copy_(self.arena.peer(r, s))
        self.arena.barrier()  # nobody overwrites its arena (next call) while a peer still reads it

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.no_grad():
                self._run()
        # a replay rewrites the static outputs without bumping their _version: derived-data caches keyed on tensor
        # versions (channels-last copies of the volumes, the TV value) must not outlive it
        from . import projector
        projector.clear_caches()
        return self.volumes, self.mask_volumes
