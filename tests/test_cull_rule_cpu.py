"""The frustum-culling rule of K1's row-group kernel (csrc/volume_agg.cu: cull_planes) restated in numpy fp32 and
checked on the CPU against the C oracle's per-view validity: a (tile, view) the rule culls must not contain a single
valid voxel, for the bench scene and for camera arrangements that stress the rule.  (The CUDA implementation itself
is pinned bit for bit by tests/test_volume_gpu.py::test_rowgroup_kernel_culling_is_bit_identical.)"""
import numpy as np
import pytest
import torch

from gens_b200.synthetic import make_scene
from oracle import c_oracle

f32 = np.float32


def cull_planes(w, k, X, Y, Z, hx, hy):
    """Bit mask of the frustum planes the points are outside of with margin; mirrors the device function."""
    c, s = [], []
    for r in range(3):
        c.append(f32(w[r, 0]) * X + f32(w[r, 1]) * Y + f32(w[r, 2]) * Z + f32(w[r, 3]))
        s.append(np.abs(f32(w[r, 0]) * X) + np.abs(f32(w[r, 1]) * Y) + np.abs(f32(w[r, 2]) * Z) + np.abs(f32(w[r, 3])))
    depth = c[2]
    img0 = f32(k[0, 0]) * c[0] + f32(k[0, 2]) * c[2]
    s0 = abs(f32(k[0, 0])) * s[0] + abs(f32(k[0, 2])) * s[2]
    img1 = f32(k[1, 1]) * c[1] + f32(k[1, 2]) * c[2]
    s1 = abs(f32(k[1, 1])) * s[1] + abs(f32(k[1, 2])) * s[2]
    tau, m = f32(1e-4), f32(1e-3)
    lx, rx, ly, ry = m * hx, (f32(2) + m) * hx, m * hy, (f32(2) + m) * hy
    bits = (depth < -tau * s[2]).astype(np.uint32)
    bits |= (-img0 - lx * depth > tau * (s0 + lx * s[2])).astype(np.uint32) * 2
    bits |= (img0 - rx * depth > tau * (s0 + rx * s[2])).astype(np.uint32) * 4
    bits |= (-img1 - ly * depth > tau * (s1 + ly * s[2])).astype(np.uint32) * 8
    bits |= (img1 - ry * depth > tau * (s1 + ry * s[2])).astype(np.uint32) * 16
    return bits


def _inside(c):
    c = c.clone(); c[:, :3, 3] *= 0.3; return c


def _turned(c):
    c = c.clone()
    c[1, :3, :3] = c[1, :3, :3] @ torch.tensor([[-1.0, 0, 0], [0, 1, 0], [0, 0, -1.0]])
    c[2, :3, :3] = c[2, :3, :3] @ torch.tensor([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
    return c


def _tangent(c):
    c = c.clone(); c[:, 0, 3] += 1.7; return c


@pytest.mark.parametrize("name,mod", [("bench", None), ("inside", _inside), ("turned", _turned), ("tangent", _tangent)])
def test_culled_tiles_hold_no_valid_voxel(name, mod):
    h, w, nv, d, zs = 240, 320, 3, 64, 64
    sc = make_scene(h, w, nv, seed=11, with_images=False)
    c2ws = sc.c2ws if mod is None else mod(sc.c2ws)
    w2c = torch.inverse(c2ws).numpy().astype(f32)
    k = sc.intrs.numpy().astype(f32)
    g = torch.linspace(-1, 1, d).numpy().astype(f32)
    hx, hy = f32((w - 1) / 2), f32((h - 1) / 2)
    valid = c_oracle.volume_agg(np.zeros((nv, 4, h, w), f32), w2c, k, g, div_mode=c_oracle.DIV_RECIP, debug=True)[-1]
    valid = np.asarray(valid).reshape(nv, d, d, d).astype(bool)
    culled = 0
    for v in range(nv):
        bits = None
        for yc in (g[0::8], g[7::8]):          # the four corners of every tile of 8 rows x 64 voxels, all x planes
            for zc in (g[0::zs], g[zs - 1::zs]):
                b = cull_planes(w2c[v], k[v], g[:, None, None], yc[None, :, None], zc[None, None, :], hx, hy)
                bits = b if bits is None else bits & b
        dead = bits != 0
        seen = valid[v].reshape(d, d // 8, 8, d // zs, zs).any(4).any(2)
        assert not (dead & seen).any(), f"{name}: view {v}: {(dead & seen).sum()} culled tiles hold valid voxels"
        culled += int(dead.sum())
    if name != "bench":
        assert culled > 0, f"{name}: the rule never fired, the case tests nothing"
