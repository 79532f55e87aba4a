"""ATen-op restatement of the GenS hot path -- TEST INFRASTRUCTURE and CPU/GPU baseline arm.

The reference's hot path is thin Python over PyTorch ATen ops (the third-party dependency
that holds the arithmetic: torch==1.13.1 pinned by the reference, torch 2.11 here).  This
module restates the same op sequence, device-agnostic, so that
  * on the GPU box it gives the reference-path numerics of ATen's CUDA kernels (mask parity
    in GENS_DIV_RECIP mode) and the "reference ops on the same B200" timing, and
  * on host cores it is the multi-threaded CPU baseline of bench.py (`--impl reference`).
It is pinned against the golden fixtures produced by the real reference
(tests/test_oracle_golden.py).  Never imported by gens_b200/.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def voxel_centres(d: int, like: torch.Tensor) -> torch.Tensor:
    """(3, d^3) world coordinates; flat index a*d*d + b*d + c <-> (g[a], g[b], g[c]).
    Reference: volume.py:28-31."""
    g = torch.linspace(-1, 1, d).type_as(like)
    return torch.stack(torch.meshgrid(g, g, g, indexing="ij")).reshape(3, -1)


def project_voxels(intrs, c2ws, scale: int, d: int, hw):
    """Normalised sampling grid (nv, d^3, 2) and per-view validity (nv, d^3).
    Reference: volume.py:24-25, :32-43."""
    h, w = hw
    nv = intrs.shape[0]
    k = intrs.clone()
    k[:, :2] *= 0.5 ** scale
    xyz = voxel_centres(d, k)
    homo = torch.cat([xyz, torch.ones_like(xyz[:1])], 0).unsqueeze(0).repeat(nv, 1, 1)
    img = torch.matmul(k, torch.matmul(torch.inverse(c2ws), homo))[:, :3]
    xy = img[:, :2] / (img[:, 2:] + 1e-8)
    nx = xy[:, 0] / ((w - 1) / 2) - 1
    ny = xy[:, 1] / ((h - 1) / 2) - 1
    valid = (nx.abs() <= 1) & (ny.abs() <= 1) & (img[:, 2] > 0)
    return torch.stack([nx, ny], -1), valid


def agg_mean_var_scale(feat, intrs, c2ws, scale: int, d: int, min_vis_view: int = 1):
    """One scale of Volume.agg_mean_var (reference volume.py:21-58) -> (1,2c,d,d,d), (1,1,d,d,d)."""
    nv, c, h, w = feat.shape
    with torch.no_grad():
        grid, valid = project_voxels(intrs, c2ws, scale, d, (h, w))
    m = valid.unsqueeze(1)
    warped = F.grid_sample(feat, grid.unsqueeze(1), padding_mode="zeros", align_corners=True).squeeze(2) * m
    total, total_sq, n = warped.sum(0), (warped ** 2).sum(0), m.sum(0)
    n_safe = torch.where(n <= 0, torch.ones_like(n) * 1e-8, n)
    mean = total / n_safe
    var = total_sq / n_safe - mean ** 2
    vol = torch.cat([mean, var], 0).reshape(1, 2 * c, d, d, d)
    return vol, (n > min_vis_view).float().reshape(1, 1, d, d, d)


def agg_mean_var(features, intrs, c2ws, dims, min_vis_view: int = 1):
    out = [agg_mean_var_scale(features[i], intrs, c2ws, i, d, min_vis_view) for i, d in enumerate(dims)]
    return [o[0] for o in out], [o[1] for o in out]


def corner_indices(grid, hw):
    """floor() corner of the bilinear footprint, as ATen's align_corners=True un-normalise gives it."""
    h, w = hw
    ix = ((grid[..., 0] + 1) / 2) * (w - 1)
    iy = ((grid[..., 1] + 1) / 2) * (h - 1)
    return torch.floor(ix).to(torch.int32), torch.floor(iy).to(torch.int32)


# ---- lookup_volume (reference projector.py:217-245) -------------------------------------------
def trilinear_dd(vol, pts):
    """Double-differentiable trilinear look-up of a (1,C,D,D,D) volume at (n,3) points, built from
    elementary torch ops (ATen's grid_sampler_3d_backward has no derivative of its own; the reference
    needs its CUDA-only grad2 kernel for that).  p=(p0,p1,p2) addresses tensor dims (2,3,4);
    align_corners=True, zeros padding.  Returns (n,C)."""
    _, c, d, _, _ = vol.shape
    u = ((pts + 1) / 2) * (d - 1)
    i0 = torch.floor(u).detach()
    t = u - i0
    flat = vol.reshape(c, -1)
    out = 0
    for da in (0, 1):
        for db in (0, 1):
            for dc in (0, 1):
                a, b, cc = i0[:, 0] + da, i0[:, 1] + db, i0[:, 2] + dc
                w = (t[:, 0] if da else 1 - t[:, 0]) * (t[:, 1] if db else 1 - t[:, 1]) * (t[:, 2] if dc else 1 - t[:, 2])
                ok = (a >= 0) & (a < d) & (b >= 0) & (b < d) & (cc >= 0) & (cc < d)
                idx = (a.clamp(0, d - 1) * d + b.clamp(0, d - 1)) * d + cc.clamp(0, d - 1)
                out = out + (flat[:, idx.long()] * ok.to(vol.dtype) * w).t()
    return out


def lookup_volume(pts, volumes, sample_mode="grad"):
    """ATen-op restatement of projector.lookup_volume for a list of volumes -> (n, sum C)."""
    grid = pts.reshape(1, 1, 1, -1, 3).flip(-1)
    outs = []
    for v in volumes:
        if sample_mode == "grad":
            o = F.grid_sample(v, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        else:
            o = F.grid_sample(v, grid, mode="nearest", padding_mode="zeros", align_corners=False)
        outs.append(o.reshape(v.shape[1], -1).t())
    return torch.cat(outs, -1)
