"""Harness shims that let the UNMODIFIED reference hot path import and run on CPU in the build
container.  Nothing here changes the arithmetic of what is recorded except
where the reference cannot run on CPU at all:

  * `mcubes` is absent                      -> empty stub module (only extract_geometry's last step uses it)
  * the CUDA-only second-derivative op      -> pure-torch, double-differentiable trilinear sampler with
    (cuda_gridsample.py asserts is_cuda)       ATen's conventions (align_corners=True, zeros padding)
  * hard-coded `.cuda()` (implicit_surface.py:270) -> identity on CPU
"""
from __future__ import annotations

import sys
import types

import torch

REF = "/root/reference"


def trilinear_zeros_align(inp: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """inp (1,C,D,H,W), grid (1,1,1,n,3) with (x,y,z)->(W,H,D).  Returns (1,C,1,1,n)."""
    _, c, d, h, w = inp.shape
    g = grid.reshape(-1, 3)
    ix = ((g[:, 0] + 1) / 2) * (w - 1)
    iy = ((g[:, 1] + 1) / 2) * (h - 1)
    iz = ((g[:, 2] + 1) / 2) * (d - 1)
    x0, y0, z0 = torch.floor(ix).detach(), torch.floor(iy).detach(), torch.floor(iz).detach()
    flat = inp.reshape(c, -1)
    out = 0
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xc, yc, zc = x0 + dx, y0 + dy, z0 + dz
                wx = (ix - x0) if dx else (x0 + 1 - ix)
                wy = (iy - y0) if dy else (y0 + 1 - iy)
                wz = (iz - z0) if dz else (z0 + 1 - iz)
                ok = (xc >= 0) & (xc < w) & (yc >= 0) & (yc < h) & (zc >= 0) & (zc < d)
                idx = (zc.clamp(0, d - 1) * h + yc.clamp(0, h - 1)) * w + xc.clamp(0, w - 1)
                val = flat[:, idx.long()] * ok.to(inp.dtype)
                out = out + val * (wx * wy * wz)
    return out.reshape(1, c, 1, 1, -1)


def install():
    """Make `import models.modules.implicit_surface` work from /root/reference on CPU."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "mcubes" not in sys.modules:
        sys.modules["mcubes"] = types.ModuleType("mcubes")
    name = "models.modules.grid_sample_cuda.cuda_gridsample"
    if name not in sys.modules:
        shim = types.ModuleType(name)

        def grid_sample_3d(inp, grid, padding_mode="zeros", align_corners=True):
            assert padding_mode == "zeros" and align_corners
            return trilinear_zeros_align(inp, grid)

        shim.grid_sample_3d = grid_sample_3d
        pkg = types.ModuleType("models.modules.grid_sample_cuda")
        pkg.__path__ = []
        pkg.cuda_gridsample = shim
        sys.modules["models.modules.grid_sample_cuda"] = pkg
        sys.modules[name] = shim
    if not getattr(torch.Tensor, "_gens_cpu_shim", False):
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.Tensor._gens_cpu_shim = True


REF_CONF = {
    "sdf_network": dict(d_out=129, d_in=3, d_hidden=128, n_layers=6, skip_in=[3], multires=4, bias=0.5, scale=1.0,
                        geometric_init=True, weight_norm=True, feat_channels=20),
    "color_network": dict(d_feature=20),
    "variance_network": dict(init_val=0.3),
    "render": dict(n_samples=64, n_importance=64, up_sample_steps=4, perturb=0.0),
}
