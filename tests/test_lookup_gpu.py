"""Parity of K2 (nearest masks) and K3 (multi-scale trilinear, three differentiation levels)."""
import numpy as np
import pytest
import torch

from gens_b200 import projector
from gens_b200.synthetic import make_reg_volumes
from oracle import torch_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL, ATOL = 1e-4, 1e-6
DIMS = [32, 16, 8, 4, 2]


def _golden(golden_dir):
    return np.load(f"{golden_dir}/render.npz")


def _close(a, b, scale=None, atol=ATOL):
    scale = b.abs().max().item() if scale is None else scale
    return bool(torch.all((a - b).abs() <= atol * max(scale, 1.0) + RTOL * b.abs()))


def test_nearest_masks_bit_exact(cuda_lib, golden_dir):
    g = _golden(golden_dir)
    pts = torch.from_numpy(g["lv_pts"]).to(DEV)
    masks = [torch.from_numpy(g[f"mask{i}"].astype(np.float32))[None, None].to(DEV) for i in range(5)]
    # (a) the reference's CPU run (ATen CPU flavour of the un-normalise)
    projector.ATEN_CUDA_FLAVOUR = 0
    try:
        each = projector.lookup_volume(pts, masks, "nearest")
    finally:
        projector.ATEN_CUDA_FLAVOUR = 1
    assert np.array_equal(each.cpu().numpy(), g["lv_nearest"])
    # (b) the same ATen ops on this GPU, 2M points incl. voxel-boundary ties, default flavour
    gen = torch.Generator().manual_seed(1)
    big = (torch.rand(2_000_000, 3, generator=gen) * 2.4 - 1.2)
    big[:100000] = (torch.randint(-40, 40, (100000, 3), generator=gen).float() / 32.0)  # exact half-voxel ties
    big = big.to(DEV)
    ref = torch_oracle.lookup_volume(big, masks, "nearest")
    got = projector.lookup_volume(big, masks, "nearest")
    assert torch.equal(got, ref), f"{(got != ref).sum().item()} nearest-mask mismatches vs ATen CUDA"
    assert torch.equal(projector.mask_nearest(big, masks), ref.any(dim=-1))


def test_trilinear_forward(cuda_lib, golden_dir):
    g = _golden(golden_dir)
    vols = [v.to(DEV) for v in make_reg_volumes(DIMS, seed=11)]
    pts = torch.from_numpy(g["lv_pts"]).to(DEV)
    got = projector.lookup_volume(pts, vols)
    ref = torch.from_numpy(g["lv_feat"]).to(DEV)
    assert got.shape == ref.shape and _close(got, ref)
    assert _close(got, torch_oracle.lookup_volume(pts, vols))
    single = projector.lookup_volume(pts, vols[0])
    assert torch.equal(single, got[:, :4])


def test_trilinear_first_and_second_order(cuda_lib):
    """d/dpts, d/dvolume and the backward-of-backward against autograd of a pure-torch restatement."""
    torch.manual_seed(0)
    vols = [v.to(DEV) for v in make_reg_volumes([16, 8, 4], seed=3)]
    pts = (torch.rand(5000, 3, device=DEV) * 2.3 - 1.15)
    w1 = torch.randn(5000, 12, device=DEV)
    w2 = torch.randn(5000, 3, device=DEV)

    def run(fn):
        p = pts.clone().requires_grad_(True)
        vs = [v.clone().requires_grad_(True) for v in vols]
        f = fn(p, vs)
        (gp,) = torch.autograd.grad(f, p, w1, create_graph=True)
        gv = torch.autograd.grad(f, vs, w1, retain_graph=True)
        # second order: differentiate <gp, w2> w.r.t. pts, w1-cotangent and volumes
        gf = torch.autograd.grad(f, p, w1.clone().requires_grad_(True), create_graph=True)
        second = torch.autograd.grad((gp * w2).sum(), [p] + vs, allow_unused=True)
        return f, gp, gv, second

    mine = run(lambda p, vs: projector.lookup_volume(p, vs))
    ref = run(lambda p, vs: torch.cat([torch_oracle.trilinear_dd(v, p) for v in vs], -1))
    assert _close(mine[0], ref[0])
    assert _close(mine[1], ref[1])
    for a, b in zip(mine[2], ref[2]):
        assert _close(a, b)
    for a, b in zip(mine[3], ref[3]):
        assert (a is None) == (b is None)
        if a is not None:
            assert _close(a, b)


def test_trilinear_grad_wrt_cotangent(cuda_lib):
    """gg_out: the backward-of-backward term w.r.t. the incoming gradient (used when the SDF MLP is trained
    through its own spatial gradient)."""
    torch.manual_seed(1)
    vols = [v.to(DEV) for v in make_reg_volumes([8, 4], seed=5)]
    pts = (torch.rand(2000, 3, device=DEV) * 2.1 - 1.05)
    w2 = torch.randn(2000, 3, device=DEV)

    def run(fn):
        p = pts.clone().requires_grad_(True)
        cot = torch.randn(2000, 8, device=DEV, generator=None).requires_grad_(True)
        torch.manual_seed(2)
        cot = torch.randn(2000, 8, device=DEV).requires_grad_(True)
        f = fn(p)
        (gp,) = torch.autograd.grad(f, p, cot, create_graph=True)
        (gc,) = torch.autograd.grad((gp * w2).sum(), cot)
        return gc

    a = run(lambda p: projector.lookup_volume(p, vols))
    b = run(lambda p: torch.cat([torch_oracle.trilinear_dd(v, p) for v in vols], -1))
    assert _close(a, b)


def test_pack_cache_follows_in_place_updates(cuda_lib):
    vol = make_reg_volumes([8], seed=1)[0].to(DEV)
    pts = torch.rand(100, 3, device=DEV) * 2 - 1
    a = projector.lookup_volume(pts, vol)
    vol.mul_(2.0)  # what Adam does to the fine-tuned volumes
    b = projector.lookup_volume(pts, vol)
    assert _close(b, 2 * a)


def test_empty_and_errors(cuda_lib):
    vol = make_reg_volumes([8], seed=1)[0].to(DEV)
    assert projector.lookup_volume(torch.zeros(0, 3, device=DEV), vol).shape == (0, 4)
    with pytest.raises(RuntimeError):
        projector.lookup_volume(torch.zeros(4, 3), vol)  # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        projector.lookup_volume(torch.zeros(4, 3, device=DEV), torch.zeros(1, 3, 8, 8, 8, device=DEV))


def test_lookup_feature_k6(cuda_lib, golden_dir):
    """K6 against the golden values of the reference (CPU flavour) and against the ATen op sequence on this
    GPU (default flavour): visibility masks bit-exact, sampled features / ray-difference features 1e-4;
    backward to the feature maps against autograd of the ATen ops."""
    from gens_b200.synthetic import make_scene
    from oracle.torch_oracle import CpuOps
    g = _golden(golden_dir)
    scene = make_scene(96, 128, 3, seed=11).to(DEV)
    pts = torch.from_numpy(g["sdf_pts"]).to(DEV)
    projector.ATEN_CUDA_FLAVOUR = 0
    try:
        fv, rd, mk = projector.lookup_feature(pts, scene.imgs, scene.intrs, scene.c2ws, scene.features)
    finally:
        projector.ATEN_CUDA_FLAVOUR = 1
    assert np.array_equal(mk.cpu().numpy(), g["lf_mask"])
    # the synthetic feature maps are white noise (|df/dx| ~ 1 per pixel), so the 1-ulp differences of a pixel
    # coordinate near 100 (1e-5) between two correct fp32 evaluations show up 1:1 in the samples: atol 5e-5
    assert _close(fv, torch.from_numpy(g["lf_feat"]).to(DEV), atol=5e-5)
    assert _close(rd, torch.from_numpy(g["lf_raydiff"]).to(DEV))
    # 5 views, 300k points incl. behind-camera and off-image ones, vs ATen on the GPU
    sc5 = make_scene(96, 128, 5, seed=2).to(DEV)
    gen = torch.Generator().manual_seed(4)
    big = (torch.rand(300_000, 3, generator=gen) * 4 - 2).to(DEV)
    feats = [f.clone().requires_grad_(True) for f in sc5.features]
    fv, rd, mk = projector.lookup_feature(big, sc5.imgs, sc5.intrs, sc5.c2ws, feats)
    feats_r = [f.clone().requires_grad_(True) for f in sc5.features]
    fr, rr, mr = CpuOps.lookup_feature(big, sc5.imgs, sc5.intrs, sc5.c2ws, feats_r)
    assert torch.equal(mk, mr), f"{(mk != mr).sum().item()} visibility mismatches vs ATen CUDA"
    assert _close(fv, fr, atol=5e-5) and _close(rd, rr, atol=5e-6)
    w = torch.randn_like(fv)
    ga = torch.autograd.grad(fv, feats, w)
    gb = torch.autograd.grad(fr, feats_r, w)
    for a, b in zip(ga, gb):
        assert _close(a, b, scale=b.abs().max().item() * 10)


def test_full_size_voxel_centre_properties(cuda_lib):
    """BASELINE config 2 pyramid (256..16): a trilinear look-up at a voxel centre (align_corners=True lattice)
    returns that voxel of every scale's volume, and a nearest look-up at a cell centre (align_corners=False lattice)
    returns that cell of the mask exactly -- the size-independent statement of projector.py:217-245's conventions
    (points are (x, y, z), tensor dims 2/3/4 are x/y/z)."""
    dims = [256, 128, 64, 32, 16]
    gen = torch.Generator(device=DEV).manual_seed(7)
    vols = [torch.randn(1, 4, d, d, d, device=DEV, generator=gen) for d in dims]
    masks = [(torch.rand(1, 1, d, d, d, device=DEV, generator=gen) > 0.5).float() for d in dims]
    n = 200_000
    for s, d in enumerate(dims):
        idx = torch.randint(0, d, (n, 3), device=DEV, generator=gen)
        # trilinear: lattice of align_corners=True
        pts = torch.linspace(-1, 1, d, device=DEV)[idx]
        got = projector.lookup_volume(pts, vols)[:, 4 * s:4 * s + 4]
        want = vols[s][0, :, idx[:, 0], idx[:, 1], idx[:, 2]].t()
        # the un-normalised coordinate misses the integer by up to D * 2^-23, times the jump to the next voxel of a
        # white-noise volume (a few units): a few 1e-4 at D = 256
        assert (got - want).abs().max().item() <= 1e-3, f"scale {s}: {(got - want).abs().max().item()}"
        # nearest: cell centres of align_corners=False
        ctr = (idx.float() + 0.5) / d * 2 - 1
        each = projector.lookup_volume(ctr, masks, "nearest")[:, s]
        assert torch.equal(each, masks[s][0, 0, idx[:, 0], idx[:, 1], idx[:, 2]]), f"scale {s}"
