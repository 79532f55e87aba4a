"""Pins the oracles (C restatement + ATen-op restatement) to golden vectors made by the real
reference on CPU (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import c_oracle, torch_oracle


def _load(golden_dir):
    return np.load(f"{golden_dir}/volume_agg.npz")


def test_c_oracle_volume_matches_reference_bit_exact(golden_dir):
    g = _load(golden_dir)
    w2c = torch.inverse(torch.from_numpy(g["c2ws"])).numpy()
    for i, d in enumerate(g["dims"]):
        d = int(d)
        k = torch.from_numpy(g["intrs"]).clone()
        k[:, :2] *= 0.5 ** i
        grid = torch.linspace(-1, 1, d).numpy()
        vol, msk, ix0, iy0, vm = c_oracle.volume_agg(g[f"feat{i}"], w2c, k.numpy(), grid, debug=True)
        # masks and visible-view counts: bit-exact
        assert np.array_equal(msk, g[f"mask{i}"])
        assert np.array_equal(vm, g[f"viewmask{i}"])
        assert np.array_equal(vm.sum(0).astype(np.int32), g[f"count{i}"])
        # bilinear corner indices (where the view is valid): bit-exact
        h, w = g[f"feat{i}"].shape[-2:]
        rix, riy = torch_oracle.corner_indices(torch.from_numpy(g[f"grid{i}"]), (h, w))
        sel = vm.astype(bool)
        assert np.array_equal(ix0[sel], rix.numpy().reshape(vm.shape)[sel])
        assert np.array_equal(iy0[sel], riy.numpy().reshape(vm.shape)[sel])
        # mean / variance volumes: tolerance 1e-6 + 1e-5*|ref| (observed: 0)
        ref = g[f"volume{i}"]
        assert np.all(np.abs(vol - ref) <= 1e-6 + 1e-5 * np.abs(ref))


def test_c_oracle_slab_equals_full(golden_dir):
    g = _load(golden_dir)
    w2c = torch.inverse(torch.from_numpy(g["c2ws"])).numpy()
    d = int(g["dims"][0])
    grid = torch.linspace(-1, 1, d).numpy()
    full_v, full_m = c_oracle.volume_agg(g["feat0"], w2c, g["intrs"], grid)
    part_v = np.zeros_like(full_v)
    part_m = np.zeros_like(full_m)
    for a0 in range(0, d, d // 4):
        v, m = c_oracle.volume_agg(g["feat0"], w2c, g["intrs"], grid, a0=a0, a1=a0 + d // 4)
        part_v[:, a0:a0 + d // 4] = v[:, a0:a0 + d // 4]
        part_m[a0:a0 + d // 4] = m[a0:a0 + d // 4]
    assert np.array_equal(full_v, part_v) and np.array_equal(full_m, part_m)


def test_torch_oracle_volume_matches_reference(golden_dir):
    g = _load(golden_dir)
    feats = [torch.from_numpy(g[f"feat{i}"]) for i in range(len(g["dims"]))]
    vols, masks = torch_oracle.agg_mean_var(feats, torch.from_numpy(g["intrs"]), torch.from_numpy(g["c2ws"]),
                                            [int(d) for d in g["dims"]])
    for i in range(len(feats)):
        assert np.array_equal(masks[i][0, 0].numpy(), g[f"mask{i}"])
        assert np.array_equal(vols[i][0].numpy(), g[f"volume{i}"])


def test_div_recip_differs_only_in_last_bit(golden_dir):
    """GENS_DIV_RECIP (ATen CUDA's a*(1/b)) may flip validity only where |n| is within an ulp of 1."""
    g = _load(golden_dir)
    w2c = torch.inverse(torch.from_numpy(g["c2ws"])).numpy()
    d = int(g["dims"][0])
    grid = torch.linspace(-1, 1, d).numpy()
    _, _, _, _, vm0 = c_oracle.volume_agg(g["feat0"], w2c, g["intrs"], grid, debug=True, div_mode=c_oracle.DIV_TRUE)
    _, _, _, _, vm1 = c_oracle.volume_agg(g["feat0"], w2c, g["intrs"], grid, debug=True, div_mode=c_oracle.DIV_RECIP)
    assert (vm0 != vm1).mean() < 1e-4


def _render_golden(golden_dir):
    return np.load(f"{golden_dir}/render.npz")


def _render_volumes():
    from gens_b200.synthetic import make_reg_volumes
    return make_reg_volumes([32, 16, 8, 4, 2], seed=11)


def test_oracle_lookup_volume_matches_reference(golden_dir):
    """Nearest mask look-ups: bit-exact.  Trilinear features: 1e-6 + 1e-5*|ref| (observed ~1e-7)."""
    g = _render_golden(golden_dir)
    pts = g["lv_pts"]
    vols = _render_volumes()
    for i in range(5):
        got = c_oracle.nearest(pts, g[f"mask{i}"].astype(np.float32), fused=0)
        assert np.array_equal(got, g["lv_nearest"][:, i])
        feat = c_oracle.trilinear(pts, vols[i][0].numpy())
        ref = g["lv_feat"][:, 4 * i:4 * i + 4]
        assert np.all(np.abs(feat - ref) <= 1e-6 + 1e-5 * np.abs(ref))
    tf = torch_oracle.lookup_volume(torch.from_numpy(pts), vols).numpy()
    assert np.all(np.abs(tf - g["lv_feat"]) <= 1e-6 + 1e-5 * np.abs(g["lv_feat"]))
    tn = torch_oracle.lookup_volume(torch.from_numpy(pts), [torch.from_numpy(g[f"mask{i}"].astype(np.float32))[None, None]
                                                            for i in range(5)], "nearest").numpy()
    assert np.array_equal(tn, g["lv_nearest"])
    dd = torch.cat([torch_oracle.trilinear_dd(v, torch.from_numpy(pts)) for v in vols], -1).numpy()
    assert np.all(np.abs(dd - g["lv_feat"]) <= 1e-6 + 1e-5 * np.abs(g["lv_feat"]))


def test_loss_restatement_matches_reference_loss_terms(golden_dir):
    """oracle/torch_oracle.loss_forward + compute_lncc (restating models/losses/loss.py:23-84, ncc.py:7-50) on the
    outputs the reference recorded for the training fixture reproduce the loss terms the reference's own Loss
    computed from them (tests/golden/train.npz, written by make_golden.py train)."""
    import sys
    sys.path.insert(0, golden_dir)
    from make_golden import LOSS_CONF, train_inputs
    from oracle import torch_oracle
    g = np.load(f"{golden_dir}/train.npz")
    preds = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("out/")}
    target = train_inputs()[-1]
    got = torch_oracle.loss_forward(preds, {"color": target}, LOSS_CONF)
    for k, v in got.items():
        ref = float(g["loss/" + k])
        assert abs(float(v) - ref) <= 1e-6 * max(abs(ref), 1.0), (k, float(v), ref)
