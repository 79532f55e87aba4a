"""Volume / feature look-ups of the ray marcher -- drop-ins for the reference's
models/modules/projector.py (lookup_volume :217-245, lookup_feature :294-349, surface_patch_warp
:353-419) and the autograd triple of models/modules/grid_sample_cuda/cuda_gridsample.py.

All scales of a pyramid are sampled by ONE kernel launch (csrc/sampling.cu); re-laid-out copies of
the volumes (channels-last) are owned here, keyed on the tensor object and its _version so that volumes
optimised in place during fine-tuning are re-packed when they change.
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Sequence, Tuple, Union

import torch

from . import _lib

# 1 = un-normalise nearest look-ups the way ATen's CUDA kernels do (fused multiply-subtract);
# masks are then bit-identical to the reference running on the same GPU.  0 = ATen CPU flavour.
ATEN_CUDA_FLAVOUR = 1

# id(tensor) -> (weakref to the tensor, its _version when packed, packed copy).  The weakref guards
# against id / storage reuse after the source tensor died; _version catches in-place updates.
_PACK_CACHE: Dict[int, Tuple["weakref.ref", int, torch.Tensor]] = {}
_PACK_CACHE_MAX = 32


def _as_list(volume) -> List[torch.Tensor]:
    return [volume] if isinstance(volume, torch.Tensor) else list(volume)


def packed_volume(vol: torch.Tensor) -> torch.Tensor:
    """Channels-last (D,D,D,4) copy of a (1,4,D,D,D) volume, cached per (storage, version)."""
    if vol.dim() != 5 or vol.shape[0] != 1 or vol.shape[1] != 4 or not (vol.shape[2] == vol.shape[3] == vol.shape[4]):
        raise RuntimeError(f"gens_b200 trilinear kernels need (1,4,D,D,D) volumes, got {tuple(vol.shape)}")
    key = id(vol)
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0]() is vol and hit[1] == vol._version and hit[2].device == vol.device:
        return hit[2]
    src = _lib.f32c(vol.detach())
    d = vol.shape[2]
    out = torch.empty((d, d, d, 4), device=vol.device, dtype=torch.float32)
    _lib.check(_lib.lib().gens_pack_volume(_lib.ptr(src), _lib.ptr(out), d, _lib.stream_ptr(vol.device)),
               "gens_pack_volume")
    for k in [k for k, (ref, _, _) in _PACK_CACHE.items() if ref() is None]:
        del _PACK_CACHE[k]
    if len(_PACK_CACHE) >= _PACK_CACHE_MAX:
        _PACK_CACHE.pop(next(iter(_PACK_CACHE)))
    _PACK_CACHE[key] = (weakref.ref(vol), vol._version, out)
    return out


def clear_caches():
    _PACK_CACHE.clear()


def _pts(pts: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(pts)
    return _lib.f32c(pts.reshape(-1, 3))


def _unpack_grad(g_cl: torch.Tensor) -> torch.Tensor:
    d = g_cl.shape[0]
    out = torch.empty((1, 4, d, d, d), device=g_cl.device, dtype=torch.float32)
    _lib.check(_lib.lib().gens_unpack_volume(_lib.ptr(g_cl), _lib.ptr(out), d, _lib.stream_ptr(g_cl.device)),
               "gens_unpack_volume")
    return out


class _Trilinear(torch.autograd.Function):
    """feats(n, 4S) = trilinear look-up of S volumes; three differentiation levels like the reference's
    _GridSample3dForward / _GridSample3dBackward / grad2_3d (cuda_gridsample.py:71-123)."""

    @staticmethod
    def forward(ctx, pts, *vols):
        packed = [packed_volume(v) for v in vols]
        dims = [v.shape[2] for v in vols]
        n = pts.shape[0]
        out = torch.empty((n, 4 * len(vols)), device=pts.device, dtype=torch.float32)
        pyr = _lib.make_pyramid(packed, dims)
        _lib.check(_lib.lib().gens_trilinear_fwd(_lib.ptr(pts), n, pyr, _lib.ptr(out), _lib.stream_ptr(pts.device)),
                   "gens_trilinear_fwd")
        ctx.save_for_backward(pts, *vols)
        return out

    @staticmethod
    def backward(ctx, g_out):
        pts, *vols = ctx.saved_tensors
        res = _TrilinearBackward.apply(g_out, pts, ctx.needs_input_grad[0], tuple(ctx.needs_input_grad[1:]), *vols)
        return (res[0], *res[1:])


class _TrilinearBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g_out, pts, need_pts, need_vols, *vols):
        packed = [packed_volume(v) for v in vols]
        dims = [v.shape[2] for v in vols]
        n = pts.shape[0]
        dev = pts.device
        g_out = _lib.f32c(g_out)
        g_pts = torch.empty((n, 3), device=dev, dtype=torch.float32) if need_pts else None
        g_cl = [torch.zeros_like(p) if need else None for p, need in zip(packed, need_vols)]
        pyr = _lib.make_pyramid(packed, dims)
        gpyr = _lib.make_pyramid(g_cl, dims) if any(need_vols) else None
        _lib.check(_lib.lib().gens_trilinear_bwd(_lib.ptr(pts), n, pyr, _lib.ptr(g_out),
                                                 _lib.ptr(g_pts) if need_pts else None, gpyr,
                                                 _lib.stream_ptr(dev)), "gens_trilinear_bwd")
        ctx.save_for_backward(g_out, pts, *vols)
        ctx.need_vols = need_vols
        g_vols = [(_unpack_grad(g) if g is not None else None) for g in g_cl]
        return (g_pts, *g_vols)

    @staticmethod
    def backward(ctx, gg_pts, *gg_vols):
        g_out, pts, *vols = ctx.saved_tensors
        if any(g is not None and bool(g.any()) for g in gg_vols):
            raise RuntimeError("gens_b200: second-order gradients THROUGH the volume gradient are not implemented "
                               "(the reference passes zeros here, cuda_gridsample.py:113-114)")
        if gg_pts is None:
            return (None, None, None, None, *([None] * len(vols)))
        packed = [packed_volume(v) for v in vols]
        dims = [v.shape[2] for v in vols]
        n = pts.shape[0]
        dev = pts.device
        gg_pts = _lib.f32c(gg_pts)
        gg_out = torch.empty_like(g_out)
        g2_pts = torch.empty((n, 3), device=dev, dtype=torch.float32)
        need_vols = tuple(ctx.needs_input_grad[4:])
        g_cl = [torch.zeros_like(p) if need else None for p, need in zip(packed, need_vols)]
        pyr = _lib.make_pyramid(packed, dims)
        gpyr = _lib.make_pyramid(g_cl, dims) if any(need_vols) else None
        _lib.check(_lib.lib().gens_trilinear_bwd2(_lib.ptr(pts), n, pyr, _lib.ptr(g_out), _lib.ptr(gg_pts),
                                                  _lib.ptr(gg_out), _lib.ptr(g2_pts), gpyr, _lib.stream_ptr(dev)),
                   "gens_trilinear_bwd2")
        g_vols = [(_unpack_grad(g) if g is not None else None) for g in g_cl]
        # like the reference's grad2_3d, the results carry no grad_fn: third-order terms are dropped
        return (gg_out, g2_pts, None, None, *g_vols)


def mask_nearest(pts: torch.Tensor, masks, want_each: bool = False):
    """Nearest look-up of the mask pyramid.  Returns the (n,) bool `any` reduction every caller of the
    reference applies, or the raw (n,S) float samples when want_each."""
    masks = _as_list(masks)
    p = _pts(pts)
    n = p.shape[0]
    ms = [_lib.f32c(m) for m in masks]
    for m in ms:
        _lib.require_cuda(m)
        if m.dim() != 5 or m.shape[0] != 1 or m.shape[1] != 1:
            raise RuntimeError(f"mask volumes must be (1,1,D,D,D), got {tuple(m.shape)}")
    pyr = _lib.make_pyramid(ms, [m.shape[2] for m in ms])
    if want_each:
        each = torch.empty((n, len(ms)), device=p.device, dtype=torch.float32)
        _lib.check(_lib.lib().gens_mask_nearest(_lib.ptr(p), n, pyr, ATEN_CUDA_FLAVOUR, None, _lib.ptr(each),
                                                _lib.stream_ptr(p.device)), "gens_mask_nearest")
        return each
    any_ = torch.empty((n,), device=p.device, dtype=torch.uint8)
    _lib.check(_lib.lib().gens_mask_nearest(_lib.ptr(p), n, pyr, ATEN_CUDA_FLAVOUR, _lib.ptr(any_), None,
                                            _lib.stream_ptr(p.device)), "gens_mask_nearest")
    return any_.bool()


def lookup_volume(pts: torch.Tensor, volume: Union[torch.Tensor, Sequence[torch.Tensor]], sample_mode: str = "grad"):
    """Same contract as the reference (projector.py:217-245): (n,3) points -> (n, sum C)."""
    vols = _as_list(volume)
    p = _pts(pts)
    if sample_mode == "grad":
        return _Trilinear.apply(p, *vols)
    if sample_mode == "nearest":
        if all(v.shape[1] == 1 for v in vols):
            return mask_nearest(p, vols, want_each=True)
        raise RuntimeError("gens_b200: nearest look-up is implemented for 1-channel (mask) volumes")
    raise RuntimeError(f"gens_b200: unsupported sample_mode {sample_mode!r}")
