"""ctypes binding of oracle/libgens_oracle.so (numpy in, numpy out).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs.  Never by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DIV_TRUE, DIV_RECIP = 0, 1


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libgens_oracle.so")
    src = os.path.join(_HERE, "gens_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libgens_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def volume_agg(feat, w2c, k_stage, grid, min_vis_view=1, div_mode=DIV_TRUE, a0=0, a1=None,
               debug=False):
    """One scale of Volume.agg_mean_var.  feat (nv,C,H,W) -> volume (2C,D,D,D), mask (D,D,D)."""
    feat, w2c, k_stage, grid = _f32(feat), _f32(w2c), _f32(k_stage), _f32(grid)
    nv, c, h, w = feat.shape
    d = grid.shape[0]
    a1 = d if a1 is None else a1
    vol = np.zeros((2 * c, d, d, d), np.float32)
    msk = np.zeros((d, d, d), np.float32)
    ix0 = iy0 = vm = None
    if debug:
        ix0 = np.zeros((nv, d, d, d), np.int32)
        iy0 = np.zeros((nv, d, d, d), np.int32)
        vm = np.zeros((nv, d, d, d), np.uint8)
    rc = lib().gens_oracle_volume_agg(
        _p(feat), nv, c, h, w, _p(w2c), _p(k_stage), _p(grid), d, a0, a1, int(min_vis_view),
        int(div_mode), _p(vol), _p(msk), _p(ix0), _p(iy0), _p(vm))
    if rc != 0:
        raise RuntimeError(f"gens_oracle_volume_agg failed: {rc}")
    if debug:
        return vol, msk, ix0, iy0, vm
    return vol, msk


def nearest(pts, vol, fused=0):
    """F.grid_sample(mode='nearest', align_corners=False) of a (D,D,D) volume at (n,3) points."""
    pts, vol = _f32(pts), _f32(vol)
    n, d = pts.shape[0], vol.shape[-1]
    out = np.zeros(n, np.float32)
    lib().gens_oracle_nearest(_p(pts), ctypes.c_long(n), _p(vol), d, int(fused), _p(out))
    return out


def trilinear(pts, vol):
    """grid_sample 3-D (bilinear, zeros, align_corners=True) of a (C,D,D,D) volume -> (n,C)."""
    pts, vol = _f32(pts), _f32(vol)
    n, c, d = pts.shape[0], vol.shape[0], vol.shape[-1]
    out = np.zeros((n, c), np.float32)
    lib().gens_oracle_trilinear(_p(pts), ctypes.c_long(n), _p(vol), c, d, _p(out))
    return out
