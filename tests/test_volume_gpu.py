"""Parity of K1 (fused volume aggregation) through the C ABI, on the B200."""
import ctypes

import numpy as np
import pytest
import torch

from gens_b200 import _lib
from gens_b200.synthetic import make_scene
from gens_b200.volume import Volume, agg_mean_var_scale, stage_cameras, pack_feature_maps
from oracle import c_oracle, torch_oracle

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
# fp32 tolerance of north_star: max rel err 1e-4 (+ an absolute floor for cancelling variances)
RTOL, ATOL = 1e-4, 1e-6


def _run_k1(feat, w2c, k, d, div_mode, min_vis_view=1):
    """Direct C-ABI call of K1 + its projection view on host-prepared camera matrices.  Every device
    tensor is held in a local until the synchronising .cpu() so the allocator cannot recycle it."""
    feat_d = pack_feature_maps(feat.to(DEV))
    nv, _, h, w = feat.shape
    w2c_d, k_d, grid_d = w2c.to(DEV).contiguous(), k.to(DEV).contiguous(), torch.linspace(-1, 1, d).to(DEV)
    vol = torch.empty((8, d, d, d), device=DEV)
    msk = torch.empty((d, d, d), device=DEV)
    ix0 = torch.empty((nv, d, d, d), dtype=torch.int32, device=DEV)
    iy0 = torch.empty_like(ix0)
    valid = torch.empty((nv, d, d, d), dtype=torch.uint8, device=DEV)
    L = _lib.lib()
    _lib.check(L.gens_volume_agg_fwd(_lib.ptr(feat_d), nv, h, w, _lib.ptr(w2c_d), _lib.ptr(k_d), 1.0, _lib.ptr(grid_d),
                                     d, 0, d, 0, d ** 3, min_vis_view, div_mode, _lib.ptr(vol), _lib.ptr(msk),
                                     _lib.stream_ptr()), "gens_volume_agg_fwd")
    _lib.check(L.gens_volume_project_debug(nv, h, w, _lib.ptr(w2c_d), _lib.ptr(k_d), 1.0, _lib.ptr(grid_d), d, div_mode,
                                           _lib.ptr(ix0), _lib.ptr(iy0), _lib.ptr(valid), _lib.stream_ptr()),
               "gens_volume_project_debug")
    torch.cuda.synchronize()
    return vol.cpu().numpy(), msk.cpu().numpy(), ix0.cpu().numpy(), iy0.cpu().numpy(), valid.cpu().numpy()


def _project_debug(intrs, c2ws, scale, d, hw, div_mode):
    w2c, k = stage_cameras(intrs, c2ws, scale)
    grid = torch.linspace(-1, 1, d, device=DEV)
    nv = intrs.shape[0]
    ix0 = torch.empty((nv, d, d, d), dtype=torch.int32, device=DEV)
    iy0 = torch.empty_like(ix0)
    valid = torch.empty((nv, d, d, d), dtype=torch.uint8, device=DEV)
    _lib.check(_lib.lib().gens_volume_project_debug(nv, hw[0], hw[1], _lib.ptr(w2c), _lib.ptr(k), 1.0, _lib.ptr(grid), d,
                                                    div_mode, _lib.ptr(ix0), _lib.ptr(iy0), _lib.ptr(valid),
                                                    _lib.stream_ptr()), "project_debug")
    torch.cuda.synchronize()
    return ix0.cpu().numpy(), iy0.cpu().numpy(), valid.cpu().numpy()


def test_golden_fixture_bit_exact_vs_reference_cpu(cuda_lib, golden_dir):
    """div_mode TRUE reproduces the reference's CPU run: masks and indices bit-exact, values 1e-4."""
    g = np.load(f"{golden_dir}/volume_agg.npz")
    intrs = torch.from_numpy(g["intrs"])
    w2c = torch.inverse(torch.from_numpy(g["c2ws"]))  # camera prologue on the CPU, as in the golden run
    for i, d in enumerate(g["dims"]):
        d = int(d)
        feat = torch.from_numpy(g[f"feat{i}"])
        h, w = feat.shape[-2:]
        k = intrs.clone()
        k[:, :2] *= 0.5 ** i
        vol, msk, ix0, iy0, valid = _run_k1(feat, w2c, k, d, _lib.DIV_TRUE)
        assert np.array_equal(msk, g[f"mask{i}"])
        ref = g[f"volume{i}"]
        assert np.all(np.abs(vol - ref) <= ATOL + RTOL * np.abs(ref))
        vm = g[f"viewmask{i}"]
        assert np.array_equal(valid, vm)
        rix, riy = torch_oracle.corner_indices(torch.from_numpy(g[f"grid{i}"]), (h, w))
        sel = vm.astype(bool)
        assert np.array_equal(ix0[sel], rix.numpy().reshape(vm.shape)[sel])
        assert np.array_equal(iy0[sel], riy.numpy().reshape(vm.shape)[sel])


@pytest.mark.parametrize("nv,hw,dims", [(3, (240, 320), [128, 64, 32, 16, 8, 4]), (5, (96, 128), [48, 20, 12, 6, 3])])
def test_matches_c_oracle(cuda_lib, nv, hw, dims):
    """Seeded synthetic scenes (config 1 shape incl. a packed-kernel scale, and a ragged one with
    D % 4 != 0): CUDA == C oracle, bit for bit (masks, indices AND mean/var volumes)."""
    sc = make_scene(hw[0], hw[1], nv, seed=3, n_scales=len(dims))
    for div_mode in (_lib.DIV_TRUE, _lib.DIV_RECIP):
        for i, d in enumerate(dims):
            w2c, k = stage_cameras(sc.intrs, sc.c2ws, i)
            ovol, omsk, oix, oiy, ovm = c_oracle.volume_agg(sc.features[i].numpy(), w2c.numpy(), k.numpy(),
                                                            torch.linspace(-1, 1, d).numpy(), div_mode=div_mode,
                                                            debug=True)
            vol, msk, ix0, iy0, valid = _run_k1(sc.features[i], w2c, k, d, div_mode)
            assert np.array_equal(msk, omsk), (div_mode, d)
            assert np.array_equal(valid, ovm), (div_mode, d)
            assert np.array_equal(ix0, oix) and np.array_equal(iy0, oiy), (div_mode, d)
            assert np.all(np.abs(vol - ovol) <= ATOL + RTOL * np.abs(ovol)), (div_mode, d)
            assert np.array_equal(vol, ovol), f"values not bit-identical to the oracle (div_mode {div_mode}, D {d})"


@pytest.mark.parametrize("nv", [3, 5])
def test_shipped_path_bit_identical_to_c_oracle_at_baseline_sizes(cuda_lib, nv):
    """BASELINE config 2 (480x640, 3 views) and config 3/4 shape (5 views) at the FULL volume sizes, through the
    public API (Volume.agg_mean_var -> gens_volume_build: pose inverse in the pack launch, the culling row-group
    kernel at 256^3, the packed kernel below): visibility masks AND mean/variance volumes of every scale are
    np.array_equal to the C oracle (oracle/gens_oracle.c, restating volume.py:21-58) fed with the same camera
    matrices.  The 256^3 x nv oracle pass takes a few seconds of OpenMP C."""
    host = make_scene(480, 640, nv, seed=0, with_images=False)
    sc = host.to(DEV)
    dims = [256, 128, 64, 32, 16]
    vols, masks = Volume(volume_dims=dims).agg_mean_var(sc.features, sc.intrs, sc.c2ws)
    torch.cuda.synchronize()
    w2c = _lib.invert_poses(sc.c2ws).cpu().numpy()  # bit-identical to torch.inverse on this device (tested below)
    for i, d in enumerate(dims):
        k = host.intrs.clone()
        k[:, :2] *= 0.5 ** i
        ovol, omsk = c_oracle.volume_agg(host.features[i].numpy(), w2c, k.numpy(), torch.linspace(-1, 1, d).numpy(),
                                         div_mode=c_oracle.DIV_RECIP)
        got_m = masks[i][0, 0].cpu().numpy()
        assert np.array_equal(got_m, omsk), f"nv={nv} D={d}: {(got_m != omsk).sum()} mask flips"
        got_v = vols[i][0].cpu().numpy()
        assert np.array_equal(got_v, ovol), f"nv={nv} D={d}: {(got_v != ovol).sum()} values differ from the oracle"
        del got_v, ovol


def test_public_api_matches_aten_ops_on_gpu(cuda_lib):
    """Volume.agg_mean_var (default DIV_RECIP) vs the same ATen op sequence the reference would run on
    this GPU: visibility masks bit-exact, volumes within tolerance."""
    sc = make_scene(240, 320, 3, seed=1).to(DEV)
    dims = [64, 32, 16, 8, 4]
    vols, masks = Volume(volume_dims=dims).agg_mean_var(sc.features, sc.intrs, sc.c2ws)
    rvols, rmasks = torch_oracle.agg_mean_var(sc.features, sc.intrs, sc.c2ws, dims)
    for i in range(5):
        assert vols[i].shape == rvols[i].shape and masks[i].shape == rmasks[i].shape
        assert torch.equal(masks[i], rmasks[i]), f"scale {i}: {(masks[i] != rmasks[i]).sum().item()} mask flips"
        assert torch.all((vols[i] - rvols[i]).abs() <= ATOL + RTOL * rvols[i].abs())


def test_full_size_properties(cuda_lib):
    """BASELINE config 2 sizes (480x640, 3 views, 256..16): slab builds are bit-identical to the full
    build; mask == (count of valid views > 1); invalid voxels are exactly zero."""
    sc = make_scene(480, 640, 3, seed=0).to(DEV)
    dims = [256, 128, 64, 32, 16]
    vols, masks = Volume(volume_dims=dims).agg_mean_var(sc.features, sc.intrs, sc.c2ws)
    for i, d in enumerate(dims):
        parts = [agg_mean_var_scale(sc.features[i], sc.intrs, sc.c2ws, i, d, 1, (a, a + d // 8))
                 for a in range(0, d, d // 8)]
        assert torch.equal(torch.cat([p[0] for p in parts], 2), vols[i])
        assert torch.equal(torch.cat([p[1] for p in parts], 2), masks[i])
        _, _, valid = _project_debug(sc.intrs, sc.c2ws, i, d, sc.features[i].shape[-2:], _lib.DIV_RECIP)
        cnt = valid.sum(0)
        assert np.array_equal(masks[i][0, 0].cpu().numpy(), (cnt > 1).astype(np.float32))
        dead = torch.from_numpy(cnt == 0).to(DEV)
        assert vols[i][0][:, dead].abs().max().item() == 0.0
        fill = masks[i].mean().item()
        assert 0.15 < fill < 0.6, fill
    # exact homogeneity at full size: features x 4 -> means x 4, variances x 16, masks unchanged, bit for bit
    vols4, masks4 = Volume(volume_dims=dims).agg_mean_var([f * 4 for f in sc.features], sc.intrs, sc.c2ws)
    for i in range(len(dims)):
        assert torch.equal(masks4[i], masks[i])
        assert torch.equal(vols4[i][:, :4], vols[i][:, :4] * 4) and torch.equal(vols4[i][:, 4:], vols[i][:, 4:] * 16)
    del vols4, masks4
    # largest scale against the ATen op sequence on the same device (64^3 sub-sample to bound memory)
    rvol, rmask = torch_oracle.agg_mean_var_scale(sc.features[2], sc.intrs, sc.c2ws, 2, 64)
    assert torch.equal(rmask, masks[2])
    assert torch.all((vols[2] - rvol).abs() <= ATOL + RTOL * rvol.abs())


def test_backward_matches_autograd_of_aten_ops(cuda_lib):
    sc = make_scene(96, 128, 3, seed=5).to(DEV)
    d = 32
    feat = sc.features[0].clone().requires_grad_(True)
    vol, _ = agg_mean_var_scale(feat, sc.intrs, sc.c2ws, 0, d)
    gout = torch.randn_like(vol)
    (gfeat,) = torch.autograd.grad(vol, feat, gout)
    feat2 = sc.features[0].clone().requires_grad_(True)
    rvol, _ = torch_oracle.agg_mean_var_scale(feat2, sc.intrs, sc.c2ws, 0, d)
    (rg,) = torch.autograd.grad(rvol, feat2, gout)
    scale = rg.abs().max().item()
    assert torch.all((gfeat - rg).abs() <= 1e-5 * scale + 1e-4 * rg.abs())


def test_rejects_cpu_tensors_and_bad_channels(cuda_lib):
    sc = make_scene(96, 128, 3, seed=0)
    with pytest.raises(RuntimeError):
        Volume(volume_dims=[8]).agg_mean_var(sc.features, sc.intrs, sc.c2ws)
    with pytest.raises(RuntimeError):
        pack_feature_maps(torch.zeros(1, 3, 4, 4, device=DEV))


def test_exact_division_shortcuts(cuda_lib):
    """K1's two division shortcuts are bit-identical to div.rn.f32: s/n for EVERY fp32 s and n = 1..16,
    and the shared-reciprocal projection division on 2^32 random operand pairs."""
    out = torch.zeros(2, dtype=torch.int64, device=DEV)
    _lib.check(_lib.lib().gens_selftest_division(16, 1 << 32, _lib.ptr(out), _lib.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    assert out.tolist() == [0, 0], out.tolist()


def test_public_api_backward_all_scales(cuda_lib):
    """Gradients through Volume.agg_mean_var (one autograd node for all scales) match autograd of the
    ATen op sequence for every feature map."""
    sc = make_scene(96, 128, 3, seed=9).to(DEV)
    dims = [32, 16, 8]
    feats = [f.clone().requires_grad_(True) for f in sc.features[:3]]
    vols, _ = Volume(volume_dims=dims).agg_mean_var(feats, sc.intrs, sc.c2ws)
    gouts = [torch.randn_like(v) for v in vols]
    grads = torch.autograd.grad(vols, feats, gouts)
    feats2 = [f.clone().requires_grad_(True) for f in sc.features[:3]]
    rvols, _ = torch_oracle.agg_mean_var(feats2, sc.intrs, sc.c2ws, dims)
    rgrads = torch.autograd.grad(rvols, feats2, gouts)
    for g, rg in zip(grads, rgrads):
        scale = rg.abs().max().item()
        assert torch.all((g - rg).abs() <= 1e-5 * scale + 1e-4 * rg.abs())


def test_pose_inverse_is_bit_identical_to_torch(cuda_lib):
    """gens_invert_poses (also folded into the pack launch of every build) must return the very bits of the
    reference's torch.inverse(c2ws) on this GPU: the matrices decide voxel validity at frustum borders."""
    from gens_b200 import _lib
    from gens_b200.volume import pack_feature_pyramid
    g = torch.Generator().manual_seed(17)
    n = 50000
    q, r = torch.linalg.qr(torch.randn(n, 3, 3, generator=g, dtype=torch.float64))
    q = q * torch.sign(torch.diagonal(r, dim1=1, dim2=2))[:, None, :]
    c2w = torch.zeros(n, 4, 4, dtype=torch.float64)
    c2w[:, :3, :3] = q
    c2w[:, :3, 3] = torch.randn(n, 3, generator=g, dtype=torch.float64) * 3
    c2w[:, 3, 3] = 1
    # a quarter of them: near-identity rotations (DTU-like rigs), where pivoting never permutes
    small = torch.randn(n // 4, 3, generator=g, dtype=torch.float64) * 0.2
    skew = torch.zeros(n // 4, 3, 3, dtype=torch.float64)
    skew[:, 0, 1], skew[:, 0, 2], skew[:, 1, 2] = -small[:, 2], small[:, 1], -small[:, 0]
    skew = skew - skew.transpose(1, 2)
    c2w[: n // 4, :3, :3] = torch.linalg.matrix_exp(skew)
    scenes = torch.cat([make_scene(480, 640, nv, seed=s, with_images=False).c2ws for nv in (3, 5) for s in (0, 1, 2)])
    c2w = torch.cat([scenes, c2w.float()]).to(DEV)
    ref = torch.inverse(c2w)
    got = _lib.invert_poses(c2w)
    assert torch.equal(got, ref), f"{int((got != ref).any(-1).any(-1).sum())} of {c2w.shape[0]} matrices differ"
    assert torch.equal(_lib.invert_poses(c2w[:5]), torch.linalg.inv_ex(c2w[:5])[0])
    # the copy folded into the pack launch
    feats = [torch.randn(3, 4, 12, 16, device=DEV)]
    _, inv = pack_feature_pyramid(feats, c2w[:3])
    assert torch.equal(inv, ref[:3])
    assert _lib.invert_poses(c2w[:0]).shape == (0, 4, 4)
    with pytest.raises(RuntimeError):
        _lib.invert_poses(c2w.cpu())
    # informative only: general (non-rigid) matrices
    rnd = torch.randn(20000, 4, 4, generator=g).to(DEV)
    share = float((_lib.invert_poses(rnd) == torch.inverse(rnd)).all(-1).all(-1).float().mean())
    print(f"general 4x4 matrices reproduced bit for bit: {share:.4f}")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one node (NVLink peer mappings)")
def test_fused_slab_exchange_two_gpus():
    """K1 storing its slab into both ranks' final tensors over peer mappings: bit-identical to the 1-GPU build
    (tools/check_fused_slabs.py under torchrun, world size 2)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", "tools/check_fused_slabs.py"],
                       cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("bit-identical to the 1-GPU build: True") == 2, r.stdout[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one node")
def test_slab_regulariser_handoff_two_gpus():
    """K1 slabs -> slab-parallel RegNetwork -> gather of the 4-channel results, on 2 ranks at the config-2
    sizes, against the whole-volume pipeline on every rank (bench.py's `regularise` leg carries the check), through
    NCCL messages and through NVLink peer memory + CUDA graph (reg_network.PeerSlabRegulariser)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", "bench.py", "--gpus", "2", "--steps", "2",
                        "--warmup", "3", "--no-render", "--no-lattice", "--no-train", "--no-cpu"],
                       cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    reg = line["regularise"]
    assert reg["verified"]["as_accurate_as_the_whole_volume_pipeline"] is True, reg
    assert reg["peer_memory"] is not None and reg["peer_memory"]["as_accurate_as_the_whole_volume_pipeline"] is True, reg
    assert line["verified"]["slabs_bit_identical"] is True
    assert reg["gathered_bytes_per_gpu"] * 9 == reg["gathered_bytes_if_volumes_were_exchanged_first"] * 5


def _k1_variant(variant, feat_d, nv, h, w, w2c_d, k_d, grid_d, d, a0=0, a1=None, min_vis_view=1, const_cams=False):
    """One K1 launch under a tuning variant; const_cams stages the cameras in the constant bank first."""
    from gens_b200.volume import agg_scale_into, stage_camera_slots
    a1 = d if a1 is None else a1
    planes = a1 - a0
    vol = torch.full((8, planes, d, d), float("nan"), device=DEV)
    msk = torch.full((planes, d, d), float("nan"), device=DEV)
    L = _lib.lib()
    L.gens_debug_set_variant(variant)
    try:
        slot = stage_camera_slots(w2c_d, k_d, [1.0])[0] if const_cams else 0
        assert slot > 0 or not const_cams or nv > 8
        agg_scale_into(feat_d, (h, w), w2c_d, k_d, 1.0, grid_d, d, vol, msk, (a0, a1), min_vis_view, _lib.DIV_RECIP, slot)
        torch.cuda.synchronize()
    finally:
        L.gens_debug_set_variant(0)
    return vol, msk


def _camera_cases():
    """The bench scene plus cameras that stress the frustum culling: inside the volume, looking away,
    rolled by 90 degrees, nearly tangent to the volume faces."""
    def scaled(f):
        def mod(c):
            c = c.clone(); c[:, :3, 3] *= f; return c
        return mod

    def turned(c):
        c = c.clone()
        c[1, :3, :3] = c[1, :3, :3] @ torch.tensor([[-1.0, 0, 0], [0, 1, 0], [0, 0, -1.0]])  # view 1 looks away
        r = torch.tensor([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
        c[2, :3, :3] = c[2, :3, :3] @ r  # view 2 rolled
        return c

    def tangent(c):
        c = c.clone(); c[:, 0, 3] += 1.7; return c
    return [("bench", None), ("inside", scaled(0.3)), ("centre", scaled(0.0)), ("turned", turned), ("tangent", tangent)]


@pytest.mark.parametrize("d,hw,nv", [(64, (96, 128), 3), (128, (240, 320), 3), (256, (480, 640), 5)])
def test_rowgroup_kernel_culling_is_bit_identical(cuda_lib, d, hw, nv):
    """The row-group kernel (conservative frustum culling per tile and view; the default at D >= 256, variant 20
    of the tuning knob at any D % 64 == 0) writes exactly what the packed kernel (variant 10) and the same
    kernel without culling (25) write, for camera arrangements that stress the culling, slab builds and
    min_vis_view = 0 included; at D = 64 all of them are also bit-identical to the C oracle."""
    for name, mod in _camera_cases():
        sc = make_scene(hw[0], hw[1], nv, seed=11, with_images=False)
        c2ws = sc.c2ws if mod is None else mod(sc.c2ws)
        w2c, k = stage_cameras(sc.intrs, c2ws, 0)
        feat_d = pack_feature_maps(sc.features[0].to(DEV))
        w2c_d, k_d, grid_d = w2c.to(DEV).contiguous(), k.to(DEV).contiguous(), torch.linspace(-1, 1, d).to(DEV)
        args = (feat_d, nv, hw[0], hw[1], w2c_d, k_d, grid_d, d)
        ref_vol, ref_msk = _k1_variant(10, *args)
        assert not torch.isnan(ref_vol).any() and not torch.isnan(ref_msk).any()
        if d == 64:
            ovol, omsk = c_oracle.volume_agg(sc.features[0].numpy(), w2c.numpy(), k.numpy(),
                                             torch.linspace(-1, 1, d).numpy(), div_mode=c_oracle.DIV_RECIP)
            assert np.array_equal(ref_vol.cpu().numpy(), ovol) and np.array_equal(ref_msk.cpu().numpy(), omsk), name
        # 0 = shipped (bulk zero fill of dead tiles), 11 = STG zero fill, 20 / 25 = with / without culling at any D;
        # each also with the cameras read from the constant bank instead of shared memory
        for variant in (0, 11, 20, 25):
            for const_cams in (False, True):
                vol, msk = _k1_variant(variant, *args, const_cams=const_cams)
                assert torch.equal(vol, ref_vol), f"{name}: variant {variant} const_cams {const_cams} volumes differ"
                assert torch.equal(msk, ref_msk), f"{name}: variant {variant} const_cams {const_cams} masks differ"
        # a slab in the middle of the volume, into a slab-sized buffer, and min_vis_view = 0 / -1 (no culling)
        a0, a1 = d // 4, d // 4 + d // 8
        vol, msk = _k1_variant(20, *args, a0=a0, a1=a1)
        assert torch.equal(vol, ref_vol[:, a0:a1]) and torch.equal(msk, ref_msk[a0:a1]), (name, "slab")
        for mvv in (0, -1):
            vol0, msk0 = _k1_variant(20, *args, min_vis_view=mvv)
            rvol0, rmsk0 = _k1_variant(10, *args, min_vis_view=mvv)
            assert torch.equal(vol0, rvol0) and torch.equal(msk0, rmsk0), (name, "min_vis_view", mvv)
