"""Volume construction -- drop-in for the reference's models/modules/volume.py.

`Volume.agg_mean_var(features, intrs, c2ws, min_vis_view=1)` keeps the reference signature and
return convention (reference volume.py:13-63) but runs ONE fused CUDA kernel per scale (K1,
csrc/volume_agg.cu) instead of ~320 ATen ops and >= 15 full-size temporaries.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib

# tensor / python-scalar division flavour (see include/gens_b200.h): the reference's masks
# differ by the last bit between its own CPU and CUDA runs; default = what the reference
# computes on the device the tensors live on, i.e. CUDA.
DEFAULT_DIV_MODE = _lib.DIV_RECIP


def to_channels_last4(feat: torch.Tensor) -> torch.Tensor:
    """(n,4,h,w) NCHW -> (n,h,w,4) with our own transpose kernel."""
    _lib.require_cuda(feat)
    feat = _lib.f32c(feat)
    n, c, h, w = feat.shape
    if c != 4:
        raise RuntimeError(f"gens_b200 volume kernels are built for 4-channel feature maps, got {c}")
    out = torch.empty((n, h, w, 4), device=feat.device, dtype=torch.float32)
    _lib.check(_lib.lib().gens_nchw4_to_nhwc4(_lib.ptr(feat), _lib.ptr(out), n, h, w,
                                              _lib.stream_ptr(feat.device)), "gens_nchw4_to_nhwc4")
    return out


def stage_cameras(intrs: torch.Tensor, c2ws: torch.Tensor, scale: int):
    """The tiny host-side prologue of the reference, kept as torch ops on the tensors' device so
    the matrices are bit-identical to the reference's (volume.py:24-25, :34)."""
    k = intrs.clone()
    k[:, :2] *= 0.5 ** scale
    return _lib.f32c(torch.inverse(c2ws)), _lib.f32c(k)


class _AggMeanVar(torch.autograd.Function):
    """One scale.  Differentiable w.r.t. the feature map only (grid is under no_grad upstream)."""

    @staticmethod
    def forward(ctx, feat, w2c, k_stage, grid, d, slab, min_vis_view, div_mode):
        a0, a1 = slab
        feat_cl = to_channels_last4(feat)
        nv, h, w, _ = feat_cl.shape
        planes = a1 - a0
        vol = torch.empty((1, 8, planes, d, d), device=feat.device, dtype=torch.float32)
        msk = torch.empty((1, 1, planes, d, d), device=feat.device, dtype=torch.float32)
        _lib.check(_lib.lib().gens_volume_agg_fwd(
            _lib.ptr(feat_cl), nv, h, w, _lib.ptr(w2c), _lib.ptr(k_stage), _lib.ptr(grid), d, a0, a1, a0,
            planes * d * d, int(min_vis_view), int(div_mode), _lib.ptr(vol), _lib.ptr(msk),
            _lib.stream_ptr(feat.device)), "gens_volume_agg_fwd")
        ctx.save_for_backward(feat_cl, w2c, k_stage, grid)
        ctx.meta = (d, a0, a1, div_mode)
        ctx.mark_non_differentiable(msk)
        return vol, msk

    @staticmethod
    def backward(ctx, g_vol, _g_msk):
        feat_cl, w2c, k_stage, grid = ctx.saved_tensors
        d, a0, a1, div_mode = ctx.meta
        nv, h, w, _ = feat_cl.shape
        g_vol = _lib.f32c(g_vol)
        g_feat = torch.zeros_like(feat_cl)
        _lib.check(_lib.lib().gens_volume_agg_bwd(
            _lib.ptr(feat_cl), nv, h, w, _lib.ptr(w2c), _lib.ptr(k_stage), _lib.ptr(grid), d, a0, a1, a0,
            (a1 - a0) * d * d, int(div_mode), _lib.ptr(g_vol), _lib.ptr(g_feat),
            _lib.stream_ptr(g_vol.device)), "gens_volume_agg_bwd")
        return g_feat.permute(0, 3, 1, 2), None, None, None, None, None, None, None


def agg_mean_var_scale(feat, intrs, c2ws, scale: int, d: int, min_vis_view: int = 1,
                       slab: Optional[Tuple[int, int]] = None, div_mode: int = DEFAULT_DIV_MODE):
    """One scale of the build; `slab=(a0,a1)` restricts it to planes of tensor dim 2."""
    _lib.require_cuda(feat, intrs, c2ws)
    w2c, k_stage = stage_cameras(intrs, c2ws, scale)
    grid = torch.linspace(-1, 1, d).type_as(k_stage)  # same call as the reference (volume.py:28)
    slab = (0, d) if slab is None else slab
    return _AggMeanVar.apply(feat, w2c, k_stage, grid, d, slab, min_vis_view, div_mode)


class Volume(nn.Module):
    """Same constructor and method as the reference's Volume (volume.py:8-13)."""

    def __init__(self, confs=None, volume_dims: Optional[Sequence[int]] = None):
        super().__init__()
        self.volume_dims = list(volume_dims) if volume_dims is not None else confs.get_list("volume_dims")
        self.div_mode = DEFAULT_DIV_MODE

    def agg_mean_var(self, features: List[torch.Tensor], intrs: torch.Tensor, c2ws: torch.Tensor,
                     min_vis_view: int = 1):
        volumes, mask_volumes = [], []
        for i, d in enumerate(self.volume_dims):
            vol, msk = agg_mean_var_scale(features[i], intrs, c2ws, i, d, min_vis_view, None, self.div_mode)
            volumes.append(vol)
            mask_volumes.append(msk)
        return volumes, mask_volumes
