"""Loss-side consumer of the hot path's patches -- drop-in for the reference's models/losses/ncc.py.

`compute_LNCC(ref_gray, src_grays)` keeps the reference signature (ncc.py:7) and is what `Loss.forward` calls on
`ref_gray_val` / `sampled_gray_val` (loss.py:36); one CUDA launch (K11, csrc/lncc.cu) replaces the two permuted
copies, three product tensors, five grouped 11x11 convolutions and ~30 element-wise ops, forward and backward.
"""
from __future__ import annotations

import torch

from . import _lib


class _LNCC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref_gray, src_grays):
        s, n, p, c = src_grays.shape
        ref = _lib.f32c(ref_gray.reshape(n, p, c))
        src = _lib.f32c(src_grays)
        dev = src.device
        score = torch.empty((n, 1), device=dev, dtype=torch.float32)
        picked = torch.empty((n, 2), device=dev, dtype=torch.int32)
        _lib.check(_lib.lib().gens_lncc_fwd(_lib.ptr(ref), _lib.ptr(src), n, s, p, c, _lib.ptr(score), None,
                                            _lib.ptr(picked), _lib.stream_ptr(dev)), "gens_lncc_fwd")
        ctx.save_for_backward(ref, src, picked)
        return score

    @staticmethod
    def backward(ctx, g_score):
        ref, src, picked = ctx.saved_tensors
        s, n, p, c = src.shape
        g = _lib.f32c(g_score.reshape(n))
        g_ref, g_src = torch.empty_like(ref), torch.empty_like(src)
        _lib.check(_lib.lib().gens_lncc_bwd(_lib.ptr(ref), _lib.ptr(src), _lib.ptr(g), _lib.ptr(picked), n, s, p, c,
                                            _lib.ptr(g_ref), _lib.ptr(g_src), _lib.stream_ptr(src.device)),
                   "gens_lncc_bwd")
        return g_ref.reshape(1, n, p, c), g_src


def compute_LNCC(ref_gray: torch.Tensor, src_grays: torch.Tensor) -> torch.Tensor:
    """ref_gray (1,B,P,C), src_grays (S,B,P,C), P = patch^2 -> (B,1): mean of the two lowest per-view scores
    mean_c clamp(1 - NCC_c^2, 0, 2), exactly the reference's compute_LNCC (ncc.py:7-50)."""
    _lib.require_cuda(ref_gray, src_grays)
    if ref_gray.dim() != 4 or src_grays.dim() != 4 or ref_gray.shape[0] != 1 or ref_gray.shape[1:] != src_grays.shape[1:]:
        raise RuntimeError(f"compute_LNCC expects (1,B,P,C) and (S,B,P,C), got {tuple(ref_gray.shape)} and "
                           f"{tuple(src_grays.shape)}")
    if src_grays.shape[0] < 2:
        raise RuntimeError("compute_LNCC needs at least two source views (the reference takes topk(k=2) over them)")
    return _LNCC.apply(ref_gray, src_grays)
