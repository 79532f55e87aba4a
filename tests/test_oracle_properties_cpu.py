"""Size-independent properties of the volume-aggregation oracle (the same ones tests/test_volume_gpu.py checks on
the CUDA path at BASELINE sizes): exact homogeneity under power-of-two scaling, slab builds = slices of the full
build, mask = (valid-view count > min_vis_view), unseen voxels exactly zero, view order only moves roundings."""
import numpy as np
import pytest
import torch

from gens_b200.synthetic import make_scene
from gens_b200.volume import stage_cameras
from oracle import c_oracle


def _scene(seed, nv=3, hw=(48, 64), d=24):
    sc = make_scene(hw[0], hw[1], nv, seed=seed, with_images=False, n_scales=1)
    w2c, k = stage_cameras(sc.intrs, sc.c2ws, 0)
    grid = torch.linspace(-1, 1, d).numpy()
    return sc.features[0].numpy(), w2c.numpy(), k.numpy(), grid


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_power_of_two_scaling_is_exact(seed):
    feat, w2c, k, grid = _scene(seed)
    v1, m1 = c_oracle.volume_agg(feat, w2c, k, grid, div_mode=c_oracle.DIV_RECIP)
    v4, m4 = c_oracle.volume_agg(feat * np.float32(4), w2c, k, grid, div_mode=c_oracle.DIV_RECIP)
    assert np.array_equal(m1, m4)
    assert np.array_equal(v4[:4], v1[:4] * np.float32(4))    # means scale by 4, variances by 16, bit for bit
    assert np.array_equal(v4[4:], v1[4:] * np.float32(16))


@pytest.mark.parametrize("seed", [0, 3])
def test_slabs_are_slices_and_masks_count_views(seed):
    feat, w2c, k, grid = _scene(seed, nv=4)
    d = grid.shape[0]
    vol, msk, _, _, valid = c_oracle.volume_agg(feat, w2c, k, grid, div_mode=c_oracle.DIV_RECIP, debug=True)
    for a0, a1 in ((0, 5), (5, 17), (17, d)):
        sv, sm = c_oracle.volume_agg(feat, w2c, k, grid, div_mode=c_oracle.DIV_RECIP, a0=a0, a1=a1)
        assert np.array_equal(sv[:, a0:a1], vol[:, a0:a1]) and np.array_equal(sm[a0:a1], msk[a0:a1])
    cnt = valid.astype(np.int32).sum(0)
    for mvv in (0, 1, 2):
        _, mm = c_oracle.volume_agg(feat, w2c, k, grid, min_vis_view=mvv, div_mode=c_oracle.DIV_RECIP)
        assert np.array_equal(mm, (cnt > mvv).astype(np.float32))
    assert np.abs(vol[:, cnt == 0]).max() == 0.0
    assert 0 < (cnt > 0).mean() < 1, "the scene must have both seen and unseen voxels for this test to mean anything"


def test_view_order_moves_only_roundings():
    feat, w2c, k, grid = _scene(5, nv=4)
    perm = [2, 0, 3, 1]
    v0, m0 = c_oracle.volume_agg(feat, w2c, k, grid, div_mode=c_oracle.DIV_RECIP)
    v1, m1 = c_oracle.volume_agg(feat[perm], w2c[perm], k[perm], grid, div_mode=c_oracle.DIV_RECIP)
    assert np.array_equal(m0, m1)
    scale = np.abs(v0).max()
    assert np.abs(v0 - v1).max() <= 1e-5 * scale
