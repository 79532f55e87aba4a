"""Volume construction -- drop-in for the reference's models/modules/volume.py.

`Volume.agg_mean_var(features, intrs, c2ws, min_vis_view=1)` keeps the reference signature and
return convention (reference volume.py:13-63) but runs ONE fused CUDA kernel per scale (K1,
csrc/volume_agg.cu), all scales enqueued by a single C call, instead of ~320 ATen ops and >= 15
full-size temporaries per scale.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib

# tensor / python-scalar division flavour (see include/gens_b200.h): the reference's masks
# differ by the last bit between its own CPU and CUDA runs; default = what the reference
# computes on the device the tensors live on, i.e. CUDA.
DEFAULT_DIV_MODE = _lib.DIV_RECIP

_GRID_CACHE: Dict[Tuple[int, str], torch.Tensor] = {}


def voxel_axis(d: int, device) -> torch.Tensor:
    """torch.linspace(-1, 1, d) on `device` -- the very call of the reference (volume.py:28), cached."""
    key = (int(d), str(device))
    g = _GRID_CACHE.get(key)
    if g is None:
        g = torch.linspace(-1, 1, d, device=device, dtype=torch.float32)
        _GRID_CACHE[key] = g
    return g


def pack_feature_maps(feat: torch.Tensor) -> torch.Tensor:
    """(n,4,h,w) NCHW -> pixel pairs (n,h+1,w,8): texel (x,y) = [f(x,y,:), f(x+1,y,:)], zero row below."""
    return pack_feature_pyramid([feat])[0]


def pack_feature_pyramid(features: Sequence[torch.Tensor], poses: Optional[torch.Tensor] = None):
    """All scales with ONE kernel launch (gens_pack_feature_maps_multi).  With `poses` (n,4,4) the same
    launch also inverts them (bit-identical to torch.inverse on CUDA) and (packed, poses_inv) is returned."""
    feats = []
    for f in features:
        _lib.require_cuda(f)
        f = _lib.f32c(f)
        if f.dim() != 4 or f.shape[1] != 4:
            raise RuntimeError("gens_b200 volume kernels are built for 4-channel (n,4,h,w) feature maps, got "
                               f"{tuple(f.shape)}")
        feats.append(f)
    n = feats[0].shape[0]
    dev = feats[0].device
    outs = [torch.empty((n, f.shape[2] + 1, f.shape[3], 8), device=dev, dtype=torch.float32) for f in feats]
    k = len(feats)
    src = (ctypes.c_void_p * k)(*[f.data_ptr() for f in feats])
    dst = (ctypes.c_void_p * k)(*[o.data_ptr() for o in outs])
    hs = (ctypes.c_int * k)(*[f.shape[2] for f in feats])
    ws = (ctypes.c_int * k)(*[f.shape[3] for f in feats])
    inv = None
    if poses is not None:
        _lib.require_cuda(poses)
        if poses.dim() != 3 or poses.shape[1:] != (4, 4):
            raise RuntimeError(f"camera poses must be (n,4,4), got {tuple(poses.shape)}")
        poses = _lib.f32c(poses)
        inv = torch.empty_like(poses)
    _lib.check(_lib.lib().gens_pack_feature_maps_multi(
        src, dst, hs, ws, k, n, _lib.ptr(poses) if inv is not None else None,
        _lib.ptr(inv) if inv is not None else None, poses.shape[0] if inv is not None else 0, _lib.stream_ptr(dev)),
        "gens_pack_feature_maps_multi")
    return outs if inv is None else (outs, inv)


def stage_cameras(intrs: torch.Tensor, c2ws: torch.Tensor, scale: int):
    """(w2c, k_stage) exactly as the reference prepares them (volume.py:24-25, :34); used by tests that
    drive the single-scale C entry with pre-scaled intrinsics (k_row_scale = 1)."""
    k = intrs.clone()
    k[:, :2] *= 0.5 ** scale
    return _lib.f32c(torch.inverse(c2ws)), _lib.f32c(k)


def stage_camera_slots(w2c: torch.Tensor, intrs: torch.Tensor, k_row_scales: Sequence[float]) -> List[int]:
    """Camera matrices of the given scales into the constant bank (gens_stage_cameras).  Returns one slot id per
    scale for `agg_scale_into(cam_slot=...)`; 0 = the constant-bank path does not apply (more than 8 views)."""
    _lib.require_cuda(w2c, intrs)
    w2c, intrs = _lib.f32c(w2c), _lib.f32c(intrs)
    n = len(k_row_scales)
    scales = (ctypes.c_float * n)(*[float(x) for x in k_row_scales])
    slots = (ctypes.c_int * n)()
    _lib.check(_lib.lib().gens_stage_cameras(_lib.ptr(w2c), _lib.ptr(intrs), w2c.shape[0], scales, n, slots,
                                             _lib.stream_ptr(w2c.device)), "gens_stage_cameras")
    return list(slots)


def agg_scale_into(packed: torch.Tensor, hw, w2c, intrs, k_row_scale: float, grid: torch.Tensor, d: int, vol, msk,
                   slab=None, min_vis_view: int = 1, div_mode: int = DEFAULT_DIV_MODE, cam_slot: int = 0):
    """ONE K1 launch into caller-provided slab buffers vol (8,planes,D,D) / msk (planes,D,D) from already packed
    pixel-pair maps and staged cameras (tests, the bench's kernel-alone timing)."""
    a0, a1 = (0, d) if slab is None else slab
    sc = (_lib.VolumeScale * 1)()
    s = sc[0]
    s.feat_padded = packed.data_ptr()
    s.H, s.W, s.D = int(hw[0]), int(hw[1]), int(d)
    s.a0, s.a1, s.a_base = a0, a1, a0
    s.channel_stride = (a1 - a0) * d * d
    s.k_row_scale = float(k_row_scale)
    s.grid, s.volume, s.mask_volume = grid.data_ptr(), vol.data_ptr(), msk.data_ptr()
    s.n_peers, s.self_peer, s.cam_slot = 0, 0, int(cam_slot)
    _lib.check(_lib.lib().gens_volume_agg_fwd_multi(sc, 1, w2c.shape[0], _lib.ptr(w2c), _lib.ptr(intrs), int(min_vis_view),
                                                    int(div_mode), _lib.stream_ptr(packed.device)),
               "gens_volume_agg_fwd_multi")


def _check_maps(features) -> List[torch.Tensor]:
    feats = []
    for f in features:
        _lib.require_cuda(f)
        f = _lib.f32c(f)
        if f.dim() != 4 or f.shape[1] != 4:
            raise RuntimeError("gens_b200 volume kernels are built for 4-channel (n,4,h,w) feature maps, got "
                               f"{tuple(f.shape)}")
        feats.append(f)
    return feats


def _build(c2ws, intrs, dims, slabs, min_vis_view, div_mode, outs, features, peer_outs=None):
    """Every allocation first, then ONE C call (gens_volume_build): pack + pose inverse launch, then the
    aggregation launches, with no interpreter time between them.  Returns (vols, masks, packed, w2c)."""
    feats = _check_maps(features)
    dev = feats[0].device
    nv = feats[0].shape[0]
    n = len(dims)
    if c2ws.dim() != 3 or c2ws.shape[1:] != (4, 4) or c2ws.shape[0] != nv:
        raise RuntimeError(f"camera poses must be ({nv},4,4), got {tuple(c2ws.shape)}")
    # one scratch allocation: w2c, then the pixel-pair maps of every scale
    sizes = [nv * (f.shape[2] + 1) * f.shape[3] * 8 for f in feats]
    scratch = torch.empty(16 * nv + sum(sizes), device=dev, dtype=torch.float32)
    w2c = scratch[: 16 * nv].view(nv, 4, 4)
    packed, off = [], 16 * nv
    for f, sz in zip(feats, sizes):
        packed.append(scratch[off: off + sz].view(nv, f.shape[2] + 1, f.shape[3], 8))
        off += sz
    scales = (_lib.VolumeScale * n)()
    vols, masks, keep = [], [], []
    for i, d in enumerate(dims):
        a0, a1 = slabs[i]
        planes = a1 - a0
        sc = scales[i]
        if peer_outs is not None:
            # multi-GPU: the slab is stored straight into the FULL (1,8,D,D,D) / (1,1,D,D,D) tensors of every
            # rank (own + NVLink peer mappings); peer_outs[i] = (vol, mask, [vol pointers], [mask pointers], own index)
            vol, msk, vol_ptrs, msk_ptrs, self_peer = peer_outs[i]
            sc.n_peers, sc.self_peer = len(vol_ptrs), int(self_peer)
            for r, (pv, pm) in enumerate(zip(vol_ptrs, msk_ptrs)):
                sc.peer_volume[r], sc.peer_mask[r] = pv, pm
            base, stride = 0, d * d * d
        elif outs is None:
            # the 8 feature channels and the mask of a scale share one allocation (both views are contiguous)
            both = torch.empty((1, 9, planes, d, d), device=dev, dtype=torch.float32)
            vol, msk = both[:, :8], both[:, 8:]
            base, stride = a0, planes * d * d
        else:  # caller-provided slab buffers (views into the all-gather send buffer)
            vol, msk = outs[i]
            base, stride = a0, planes * d * d
        grid = voxel_axis(d, dev)
        sc.feat_padded = packed[i].data_ptr()
        sc.H, sc.W, sc.D = feats[i].shape[2], feats[i].shape[3], d
        sc.a0, sc.a1, sc.a_base = a0, a1, base
        sc.channel_stride = stride
        sc.k_row_scale = 0.5 ** i
        sc.grid, sc.volume, sc.mask_volume = grid.data_ptr(), vol.data_ptr(), msk.data_ptr()
        vols.append(vol)
        masks.append(msk)
        keep.append(grid)
    src = (ctypes.c_void_p * n)(*[f.data_ptr() for f in feats])
    dst = (ctypes.c_void_p * n)(*[p.data_ptr() for p in packed])
    hs = (ctypes.c_int * n)(*[f.shape[2] for f in feats])
    ws = (ctypes.c_int * n)(*[f.shape[3] for f in feats])
    _lib.check(_lib.lib().gens_volume_build(src, dst, hs, ws, scales, n, nv, _lib.ptr(c2ws), _lib.ptr(w2c),
                                            _lib.ptr(intrs), int(min_vis_view), int(div_mode), _lib.stream_ptr(dev)),
               "gens_volume_build")
    return vols, masks, packed, w2c


class _AggMeanVar(torch.autograd.Function):
    """All scales of one build.  Differentiable w.r.t. the feature maps only (the voxel grid is under
    no_grad in the reference, volume.py:27-44)."""

    @staticmethod
    def forward(ctx, c2ws, intrs, dims, slabs, min_vis_view, div_mode, outs, *features):
        vols, masks, packed, w2c = _build(c2ws, intrs, dims, slabs, min_vis_view, div_mode, outs, features)
        ctx.save_for_backward(w2c, intrs, *packed)
        ctx.meta = (list(dims), list(slabs), div_mode, [tuple(f.shape) for f in features])
        ctx.mark_non_differentiable(*masks)
        return (*vols, *masks)

    @staticmethod
    def backward(ctx, *grads):
        w2c, intrs, *packed = ctx.saved_tensors
        dims, slabs, div_mode, shapes = ctx.meta
        n = len(dims)
        dev = w2c.device
        out = []
        for i, d in enumerate(dims):
            g_vol = grads[i]
            if g_vol is None or not ctx.needs_input_grad[7 + i]:
                out.append(None)
                continue
            g_vol = _lib.f32c(g_vol)
            nv, _, h, w = shapes[i]
            a0, a1 = slabs[i]
            g_pad = torch.zeros((nv, h + 1, w + 1, 4), device=dev, dtype=torch.float32)
            _lib.check(_lib.lib().gens_volume_agg_bwd(
                _lib.ptr(packed[i]), nv, h, w, _lib.ptr(w2c), _lib.ptr(intrs), 0.5 ** i, _lib.ptr(voxel_axis(d, dev)),
                d, a0, a1, a0, (a1 - a0) * d * d, int(div_mode), _lib.ptr(g_vol), _lib.ptr(g_pad),
                _lib.stream_ptr(dev)), "gens_volume_agg_bwd")
            g_nchw = torch.empty(shapes[i], device=dev, dtype=torch.float32)
            _lib.check(_lib.lib().gens_unpack_feature_grads(_lib.ptr(g_pad), _lib.ptr(g_nchw), nv, h, w,
                                                            _lib.stream_ptr(dev)), "gens_unpack_feature_grads")
            out.append(g_nchw)
        return (None, None, None, None, None, None, None, *out)


def agg_mean_var(features, intrs, c2ws, dims, min_vis_view: int = 1,
                 slabs: Optional[Sequence[Tuple[int, int]]] = None, div_mode: int = DEFAULT_DIV_MODE,
                 outs=None, peer_outs=None):
    """The 5-scale build.  `slabs[i] = (a0, a1)` restricts scale i to planes of tensor dim 2 (the
    multi-GPU sharding); default = full volumes.  `outs[i] = (vol, mask)` lets the caller provide the
    (contiguous, fp32) output buffers, e.g. views into an all-gather send buffer.
    Returns (volumes, mask_volumes) as the reference."""
    _lib.require_cuda(intrs, c2ws, *features[:len(dims)])
    k = _lib.f32c(intrs)
    slabs = [(0, d) for d in dims] if slabs is None else list(slabs)
    feats = features[:len(dims)]
    wants_grad = torch.is_grad_enabled() and any(f.requires_grad for f in feats)
    if wants_grad and (peer_outs is not None or outs is not None):
        # the multi-GPU builds write through raw pointers into exchange buffers: their results carry no grad_fn,
        # so a training step would silently send zero gradient to the feature network
        raise RuntimeError("gens_b200: slab-sharded / caller-buffer volume builds are inference-only (no backward to "
                           "the feature maps); run them under torch.no_grad() or use Volume.agg_mean_var")
    if peer_outs is not None or not wants_grad:
        vols, masks, _, _ = _build(_lib.f32c(c2ws), k, tuple(dims), tuple(slabs), min_vis_view, div_mode, outs, feats,
                                   peer_outs)
        return vols, masks
    out = _AggMeanVar.apply(_lib.f32c(c2ws), k, tuple(dims), tuple(slabs), min_vis_view, div_mode, outs,
                            *features[:len(dims)])
    n = len(dims)
    return list(out[:n]), list(out[n:])


def agg_mean_var_scale(feat, intrs, c2ws, scale: int, d: int, min_vis_view: int = 1,
                       slab: Optional[Tuple[int, int]] = None, div_mode: int = DEFAULT_DIV_MODE):
    """One scale only (tests, slab experiments): same kernels, k_row_scale = 0.5**scale."""
    _lib.require_cuda(feat, intrs, c2ws)
    dev = feat.device
    w2c, k = _lib.f32c(_lib.inverse(c2ws)), _lib.f32c(intrs)
    a0, a1 = (0, d) if slab is None else slab

    class _One(torch.autograd.Function):
        @staticmethod
        def forward(ctx, f):
            packed = pack_feature_maps(f)
            nv, _, h, w = f.shape
            vol = torch.empty((1, 8, a1 - a0, d, d), device=dev, dtype=torch.float32)
            msk = torch.empty((1, 1, a1 - a0, d, d), device=dev, dtype=torch.float32)
            _lib.check(_lib.lib().gens_volume_agg_fwd(
                _lib.ptr(packed), nv, h, w, _lib.ptr(w2c), _lib.ptr(k), 0.5 ** scale, _lib.ptr(voxel_axis(d, dev)), d,
                a0, a1, a0, (a1 - a0) * d * d, int(min_vis_view), int(div_mode), _lib.ptr(vol), _lib.ptr(msk),
                _lib.stream_ptr(dev)), "gens_volume_agg_fwd")
            ctx.save_for_backward(packed)
            ctx.shape = tuple(f.shape)
            ctx.mark_non_differentiable(msk)
            return vol, msk

        @staticmethod
        def backward(ctx, g_vol, _g):
            (packed,) = ctx.saved_tensors
            nv, _, h, w = ctx.shape
            g_vol = _lib.f32c(g_vol)
            g_pad = torch.zeros((nv, h + 1, w + 1, 4), device=dev, dtype=torch.float32)
            _lib.check(_lib.lib().gens_volume_agg_bwd(
                _lib.ptr(packed), nv, h, w, _lib.ptr(w2c), _lib.ptr(k), 0.5 ** scale, _lib.ptr(voxel_axis(d, dev)), d,
                a0, a1, a0, (a1 - a0) * d * d, int(div_mode), _lib.ptr(g_vol), _lib.ptr(g_pad),
                _lib.stream_ptr(dev)), "gens_volume_agg_bwd")
            g = torch.empty(ctx.shape, device=dev, dtype=torch.float32)
            _lib.check(_lib.lib().gens_unpack_feature_grads(_lib.ptr(g_pad), _lib.ptr(g), nv, h, w,
                                                            _lib.stream_ptr(dev)), "gens_unpack_feature_grads")
            return g

    return _One.apply(feat)


class Volume(nn.Module):
    """Same constructor and method as the reference's Volume (volume.py:8-13)."""

    def __init__(self, confs=None, volume_dims: Optional[Sequence[int]] = None):
        super().__init__()
        self.volume_dims = list(volume_dims) if volume_dims is not None else confs.get_list("volume_dims")
        self.div_mode = DEFAULT_DIV_MODE

    def agg_mean_var(self, features: List[torch.Tensor], intrs: torch.Tensor, c2ws: torch.Tensor,
                     min_vis_view: int = 1):
        return agg_mean_var(features, intrs, c2ws, self.volume_dims, min_vis_view, None, self.div_mode)
