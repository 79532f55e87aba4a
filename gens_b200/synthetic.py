"""Synthetic DTU-shaped scenes (no files, no network) for tests, golden vectors and bench.

Follows the recipe of SURVEY.md section 8(d), which restates what the reference's
dataset code does to real DTU cameras:
  * pinhole intrinsics scaled from the 1600x1200 DTU calibration  (datasets/dtu.py:175-185)
  * depth range [425, 425 + 2.5*192]                               (confs/gens.conf:12-13)
  * reference camera = identity, sources relative to it            (datasets/dtu.py:316)
  * frustum-union normalisation into the unit sphere               (datasets/dtu.py:193-229, :331-341)
  * rays through pixel centres                                     (datasets/dtu.py:399-404)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

DTU_FX = 2892.33
DTU_CX = 823.2
DTU_CY = 619.07
DEPTH_MIN = 425.0
DEPTH_MAX = 425.0 + 2.5 * 192
PIVOT = 650.0


def _rot_y(deg: float) -> np.ndarray:
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def _rot_x(deg: float) -> np.ndarray:
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


@dataclass
class Scene:
    """Everything `GenS.forward` hands to the hot path (ipts keys of the reference)."""

    intrs: torch.Tensor  # (nv,4,4)
    c2ws: torch.Tensor  # (nv,4,4)
    near: torch.Tensor  # (1,1)
    far: torch.Tensor  # (1,1)
    hw: tuple
    radius: float
    features: List[torch.Tensor] = field(default_factory=list)  # 5 x (nv,4,H>>i,W>>i)
    imgs: torch.Tensor | None = None  # (nv,3,H,W)

    def rays(self, step: int = 1):
        """Rays of the reference view through every `step`-th pixel, row-major (y, x)."""
        h, w = self.hw
        intrs, c2ws = self.intrs.cpu(), self.c2ws.cpu()  # always generated on the host, then moved
        ys, xs = torch.meshgrid(
            torch.arange(0, h, step, dtype=torch.float32),
            torch.arange(0, w, step, dtype=torch.float32),
            indexing="ij",
        )
        p = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(xs.numel())], dim=-1)
        kinv = torch.inverse(intrs[0, :3, :3])
        p = p @ kinv.T
        d = p / torch.linalg.norm(p, dim=-1, keepdim=True)
        d = d @ c2ws[0, :3, :3].T
        o = c2ws[0, :3, 3].expand_as(d).contiguous()
        return o.to(self.intrs.device), d.contiguous().to(self.intrs.device)

    def to(self, device):
        return Scene(
            self.intrs.to(device), self.c2ws.to(device), self.near.to(device), self.far.to(device),
            self.hw, self.radius, [f.to(device) for f in self.features],
            None if self.imgs is None else self.imgs.to(device),
        )


def make_cameras(h: int, w: int, nv: int, factor: float = 0.8):
    """Intrinsics / normalised poses of `nv` DTU-like views (view 0 = reference)."""
    k = np.eye(4, dtype=np.float64)
    k[0, 0] = DTU_FX * w / 1600.0
    k[1, 1] = DTU_FX * h / 1200.0
    k[0, 2] = DTU_CX * w / 1600.0
    k[1, 2] = DTU_CY * h / 1200.0
    pivot = np.array([0.0, 0.0, PIVOT])
    rots = [np.eye(3), _rot_y(12.0), _rot_y(-12.0), _rot_x(12.0), _rot_x(-12.0),
            _rot_y(8.0) @ _rot_x(8.0), _rot_y(-8.0) @ _rot_x(-8.0)]
    c2w_raw = []
    for v in range(nv):
        r = rots[v % len(rots)]
        c = pivot - r @ pivot  # camera centre orbiting the pivot, still looking at it
        m = np.eye(4)
        m[:3, :3] = r
        m[:3, 3] = c
        c2w_raw.append(m)
    # union of the view frusta in world space -> centre, radius
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for m in c2w_raw:
        for z in (DEPTH_MIN, DEPTH_MAX):
            for px in (0.0, float(w)):
                for py in (0.0, float(h)):
                    pc = np.array([(px - k[0, 2]) * z / k[0, 0], (py - k[1, 2]) * z / k[1, 1], z])
                    pw = m[:3, :3] @ pc + m[:3, 3]
                    lo, hi = np.minimum(lo, pw), np.maximum(hi, pw)
    centre = (lo + hi) / 2
    radius = float((hi - lo).max() / 2 * factor)
    c2ws = []
    for m in c2w_raw:
        n = m.copy()
        n[:3, 3] = (m[:3, 3] - centre) / radius
        c2ws.append(n)
    c2ws = torch.from_numpy(np.stack(c2ws).astype(np.float32))
    intrs = torch.from_numpy(np.stack([k] * nv).astype(np.float32))
    dist = float(np.linalg.norm(c2ws[0, :3, 3].numpy()))
    near = torch.tensor([[0.95 * (dist - 1.0)]], dtype=torch.float32)
    far = torch.tensor([[1.05 * (dist + 1.0)]], dtype=torch.float32)
    return intrs, c2ws, near, far, radius


def make_scene(h: int, w: int, nv: int, seed: int = 0, n_scales: int = 5, feat_ch: int = 4,
               factor: float = 0.8, with_images: bool = True) -> Scene:
    intrs, c2ws, near, far, radius = make_cameras(h, w, nv, factor)
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(nv, feat_ch, h >> i, w >> i, generator=g) * 0.5 for i in range(n_scales)]
    imgs = torch.rand(nv, 3, h, w, generator=g) if with_images else None
    return Scene(intrs, c2ws, near, far, (h, w), radius, feats, imgs)


def make_reg_volumes(dims, seed: int = 0, ch: int = 4):
    """Stand-ins for the regularised volumes RegNetwork hands to the renderer: smooth random fields
    (low-resolution noise, trilinearly up-sampled) plus a little high-frequency noise, (1,ch,D,D,D)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed + 1000)
    vols = []
    for d in dims:
        lo_res = max(d // 4, 2)
        base = torch.randn(1, ch, lo_res, lo_res, lo_res, generator=g) * 0.5
        v = F.interpolate(base, size=(d, d, d), mode="trilinear", align_corners=True)
        v = v + 0.05 * torch.randn(1, ch, d, d, d, generator=g)
        vols.append(v.contiguous())
    return vols
