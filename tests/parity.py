"""Shared tolerance check of the parity tests.

north_star's fp32 bar: |a-b| <= atol + 1e-4*|b|, atol = atol_scale x the tensor's scale.  Quantities
that are DISCONTINUOUS in the sample position -- the gradient of a trilinear field jumps at voxel
faces, and everything it steers (true_cos -> alpha -> weights) inherits the jumps -- may miss the tight
bound on a stated small share of elements (a 1e-7 difference in a sample depth between two runs can land
on the other side of a face), but never the gross bound."""
import numpy as np
import torch

JUMPY = {"gradients": 2e-3, "normal": 2e-3, "ref_gray_val": 2e-3, "sampled_gray_val": 2e-3, "weights": 1e-2,
         "weight_sum": 1e-2, "weight_max": 1e-2}


def mismatch(name, got, ref, rtol=1e-4, atol_scale=1e-6, outlier_frac=0.0, outlier_rtol=1e-2):
    """None if `got` matches `ref`, else a message."""
    got = got.detach().float().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    if got.shape != ref.shape:
        return f"{name}: shape {got.shape} != {ref.shape}"
    scale = max(float(np.abs(ref).max()), 1.0) if ref.size else 1.0
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    bad = err > atol_scale * scale + rtol * np.abs(ref)
    gross = err > outlier_rtol * (scale + np.abs(ref))
    if bad.sum() > outlier_frac * bad.size or gross.any():
        return (f"{name}: {bad.sum()} of {bad.size} outside tolerance ({gross.sum()} gross), "
                f"max err {err.max():.3e} (scale {scale:.3e})")
    return None


def check(name, got, ref, **kw):
    msg = mismatch(name, got, ref, **kw)
    assert msg is None, msg
