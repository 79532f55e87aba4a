"""End-to-end drop-in proof against the UNMODIFIED reference on the same GPU.

The reference staged under baseline/_ref/ (baseline/setup_ref.py; its own gridsample_grad2 CUDA extension pre-built
for sm_100a) runs `GenS.forward("train")` + its own `Loss` + `backward()`, and `GenS.forward("val")`
(models/gens.py:124-157) on a synthetic scene.  Then `gens_b200.install()` patches the hot path into the very same
reference tree, a second GenS is constructed from the patched modules with the first one's state_dict, and the same
calls are repeated: every output key and every parameter gradient is compared.

Skipped when baseline/_ref is not staged (it is git-ignored; `python baseline/setup_ref.py` creates it in the
build container and gpurun ships it).
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_runtime  # noqa: E402

from gens_b200.config import Conf, gens_model_conf  # noqa: E402
from gens_b200.synthetic import make_scene  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_runtime.available(), reason=ref_runtime.why_unavailable() or "-")]

DIMS = [64, 32, 16, 8, 4]  # config-1 pyramid (the 3-D U-Net halves five times: 4^3 is its smallest input)
H, W = 96, 128


def _conf(perturb=1.0):
    c = gens_model_conf(perturb=perturb)
    d = {k: (dict(v) if isinstance(v, dict) else v) for k, v in c.items()}
    d["volume"] = {"volume_dims": DIMS}
    return Conf(d)


def _inputs(dev, nv, n_rays, seed=0, val=False):
    sc = make_scene(H, W, nv, seed=seed)
    g = torch.Generator().manual_seed(seed + 11)
    if val:
        ro, rd = sc.rays(step=4)
    else:
        ro, rd = sc.rays(step=1)
        sel = torch.randperm(ro.shape[0], generator=g)[:n_rays]
        ro, rd = ro[sel].contiguous(), rd[sel].contiguous()
    ipts = {"imgs": sc.imgs.to(dev), "intrs": sc.intrs.to(dev), "c2ws": sc.c2ws.to(dev), "rays_o": ro.to(dev),
            "rays_d": rd.to(dev), "near": sc.near.to(dev), "far": sc.far.to(dev)}
    if val:
        ipts["bound_min"] = torch.tensor([-1.0, -1.0, -1.0], device=dev)
        ipts["bound_max"] = torch.tensor([1.0, 1.0, 1.0], device=dev)
        ipts["hw"] = (H // 4, W // 4)
    else:
        ipts["pseudo_pts"] = (torch.rand(512, 3, generator=g) * 1.2 - 0.6).to(dev)
    target = torch.rand(ro.shape[0], 3, generator=g).to(dev)
    return ipts, target


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    scale = max(float(b.abs().max()), 1e-12)
    return float((a - b).abs().max()) / scale


def _run_train(ns, model, dev):
    loss_fn = ns.Loss(Conf(ref_runtime.LOSS_CONF))
    model.train()
    model.zero_grad(set_to_none=True)
    ipts, target = _inputs(dev, nv=5, n_rays=256)
    torch.manual_seed(123)
    out = model("train", ipts, cos_anneal_ratio=0.5, step=7)
    losses = loss_fn(out, {"color": target}, step=7)
    losses["loss"].backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    return ({k: v.detach().clone() for k, v in out.items()}, {k: float(v) for k, v in losses.items()}, grads)


def _run_val(ns, model, dev):
    model.eval()
    if hasattr(model.implicit_surface, "mesher"):
        model.implicit_surface.mesher = "mcubes"  # hand the lattice to the recording stub like the reference does
    ipts, _ = _inputs(dev, nv=3, n_rays=0, val=True)
    torch.manual_seed(321)
    ns.mcubes.last_u = None
    with torch.no_grad():
        out = model("val", ipts, cos_anneal_ratio=1.0)
    u = ns.mcubes.last_u
    assert u is not None and u.shape == (512, 512, 512)
    return out, np.array(u, copy=True)


@pytest.fixture(scope="module")
def arms(cuda_lib):
    """Reference results first, from the UN-PATCHED tree (install() rewrites the reference modules' globals, so
    nothing of the reference may run after it), then the same tree with the hot path patched in."""
    import gens_b200
    dev = torch.device("cuda:0")
    ns = ref_runtime.load()
    torch.manual_seed(0)
    ref = ns.GenS(_conf()).to(dev)
    ref_cls = type(ref.implicit_surface)
    assert ref_cls.__module__ == "models.modules.implicit_surface"
    assert ns.implicit_surface.lookup_volume.__module__ == "models.modules.projector"  # still the reference's own
    state = {k: v.clone() for k, v in ref.state_dict().items()}
    ref_train = _run_train(ns, ref, dev)
    ref_val = _run_val(ns, ref, dev)
    del ref
    torch.cuda.empty_cache()
    # second arm: the SAME reference tree (models/gens.py untouched) with the hot path patched in
    gens_b200.install()
    assert ns.implicit_surface.lookup_volume.__module__ == "gens_b200.projector"
    ours = ns.GenS(_conf()).to(dev)
    assert type(ours.implicit_surface) is gens_b200.ImplicitSurface and type(ours.implicit_surface) is not ref_cls
    assert type(ours.volume) is gens_b200.Volume
    ours.load_state_dict(state)
    yield ns, ref_train, ref_val, ours
    ref_runtime.purge()


def test_forward_train_loss_backward_matches_reference(arms):
    """models/gens.py:124-157 in "train" mode, then the reference's Loss (models/losses/loss.py:23-84; its
    compute_LNCC is K11 on the patched arm) and backward(): outputs, loss terms and the gradient of EVERY trainable
    parameter (2-D feature CNN, 3-D U-Net, SDF / colour / variance networks) against the un-patched reference run on
    the same GPU."""
    ns, (o_ref, l_ref, g_ref), _, ours = arms
    dev = torch.device("cuda:0")
    # The no-grad importance sampling is a chaotic amplifier (inverse CDF at inv_s up to 512, then weights with a slope
    # of inv_s ~ 20 in the depth): SDF values that agree to 1e-6 (fp32 chain) or 8e-6 (shipped 3xTF32 kernel) with the
    # reference's still move individual samples by 1e-5 ... 1e-4, and now and then one sample of one ray across a
    # voxel face or a zero crossing.  Bounds below are therefore stated per quantity as (share of elements beyond the
    # tight bound, worst element).
    o_our, l_our, g_our = _run_train(ns, ours, dev)
    assert set(o_ref) == set(o_our)
    report = {}
    for k in sorted(o_ref):
        a, b = o_our[k], o_ref[k]
        assert tuple(a.shape) == tuple(b.shape), k
        if b.dtype == torch.bool:
            report[k] = float((a != b).float().mean())
        else:
            report[k] = _rel(a, b)
    print("train outputs, max |diff| / max |ref|:", {k: f"{v:.2e}" for k, v in report.items()})
    # discrete outputs
    assert report["valid_mask"] == 0.0
    assert float((o_our["mid_inside_sphere"] != o_ref["mid_inside_sphere"]).float().mean()) <= 0.02
    # continuous per-ray outputs (fp32; the up-sampling SDF passes run on the 3xTF32 tensor-core kernel under
    # no_grad, which moves the importance samples by <= 1e-4 and everything downstream accordingly)
    def off(k, tight_rel):
        a, b = o_our[k].detach().float(), o_ref[k].detach().float()
        err = (a - b).abs() / b.abs().max().clamp_min(1e-12)
        return float((err > tight_rel).float().mean()), float(err.max())
    # per-ray / scalar outputs: all within 2e-3 of the tensor's scale, except that 1 % of the rays may have had a
    # sample flip (never beyond 5e-2)
    for k in ("color_fine", "render_depth", "weight_sum", "normal", "s_val", "tv_reg", "sparse_sdf", "pseudo_sdf",
              "gradient_error", "inside_sphere", "smooth_error", "sdf_depth"):
        share, worst_el = off(k, 2e-3)
        assert share <= 0.01 and worst_el <= 5e-2, (k, share, worst_el)
    # per-sample quantities follow the moved sample positions (and the gradient of a trilinear field jumps at voxel
    # faces): at most 1 % of the elements off by more than 2e-3 of the tensor's scale, none by more than 5e-2
    for k in ("weights", "weight_max", "gradients", "ref_gray_val", "sampled_gray_val"):
        share, worst_el = off(k, 2e-3)
        assert share <= 0.01 and worst_el <= 5e-2, (k, share, worst_el)
    print("loss terms (ours, reference):", {k: (l_our[k], l_ref[k]) for k in l_ref})
    for k in ("loss", "color_loss", "eikonal_loss", "sparse_loss", "mfc_loss", "tv_loss", "pseudo_sdf_loss"):
        assert abs(l_our[k] - l_ref[k]) <= 1e-3 * max(abs(l_ref[k]), 1e-3), (k, l_our[k], l_ref[k])
    # gradients of every parameter the reference trains
    assert set(g_ref) == set(g_our), set(g_ref) ^ set(g_our)
    # error of a tensor's gradient relative to its own largest entry, but never finer than 1e-3 of the largest
    # gradient entry of its group: parameters whose true gradient is zero (a bias feeding a normalisation layer, the
    # colour network's scalar `s`) hold cancellation noise of that size in BOTH runs
    groups = {"feature_network": [], "reg_network": [], "implicit_surface": []}
    top = {gname: max(float(g_ref[n].abs().max()) for n in g_ref if n.startswith(gname)) for gname in groups}
    worst = {}
    for n in g_ref:
        gname = next(k for k in groups if n.startswith(k))
        scale = max(float(g_ref[n].abs().max()), 1e-3 * top[gname])
        worst[n] = float((g_our[n] - g_ref[n]).abs().max()) / scale
        groups[gname].append(worst[n])
    print("train gradients, worst:", [(n, f"{v:.2e}") for n, v in sorted(worst.items(), key=lambda kv: -kv[1])[:8]])
    for gname, vals in groups.items():
        assert vals, gname
        print(f"{gname}: {len(vals)} tensors, largest |grad| {top[gname]:.2e}, median err {float(np.median(vals)):.2e}, "
              f"max err {max(vals):.2e}")
        # feature CNN and MLPs: 5e-3 median / 0.1 worst of the tensor's scale (observed 5e-4 ... 9e-4 / 0.02 ... 0.07).  The 3-D U-Net only receives gradient
        # through the volume features, which the geometric initialisation all but disconnects from the SDF (largest
        # entry ~1e-4 of the others'): what arrives is dominated by the sample-placement noise above -- 5e-2 / 0.5
        lim_med, lim_max = (5e-2, 0.5) if gname == "reg_network" else (5e-3, 0.1)
        assert float(np.median(vals)) <= lim_med, (gname, float(np.median(vals)))
        assert max(vals) <= lim_max, (gname, max(vals))


def test_forward_val_matches_reference(arms):
    """models/gens.py:124-157 in "val" mode: the 512^3 mesh-extraction lattice (implicit_surface.py:407-421, handed to
    the recording mcubes stub) and the rendered colour / depth / normal images."""
    ns, _, (o_ref, u_ref), ours = arms
    dev = torch.device("cuda:0")
    o_our, u_our = _run_val(ns, ours, dev)
    du = np.abs(u_our - u_ref)
    print(f"512^3 lattice: max |diff| {du.max():.2e}, |ref| max {np.abs(u_ref).max():.2e}")
    # 3xTF32 tensor-core value pass against the reference's fp32 cuBLAS chain
    assert du.max() <= 2e-5 + 1e-4 * np.abs(u_ref).max()
    # the device mesher (K12, what extract_geometry ships with) on the REFERENCE's 512^3 lattice: the vertex set every
    # marching-cubes implementation must produce, and a consistently oriented surface
    from gens_b200.meshing import marching_cubes
    from oracle import mc_oracle
    v, t = marching_cubes(torch.from_numpy(u_ref).to(dev), 0.0)
    assert len(t) > 0
    assert np.array_equal(v[np.lexsort((v[:, 2], v[:, 1], v[:, 0]))], mc_oracle.edge_vertices(u_ref, 0.0))
    rep = mc_oracle.mesh_report(v, t)
    print(f"device mesher on the reference lattice: {len(v)} vertices, {len(t)} triangles, {rep}")
    assert rep["oriented"] and rep["degenerate"] == 0 and rep["unused_vertices"] == 0, rep
    assert set(o_ref) == set(o_our)
    for k in ("color_fine", "img_fine", "normal_img", "sdf_depth", "render_depth"):
        a = torch.as_tensor(np.asarray(o_our[k])).float()
        b = torch.as_tensor(np.asarray(o_ref[k])).float()
        assert a.shape == b.shape, k
        err = (a - b).abs()
        scale = float(b.abs().max()) + 1e-12
        frac_bad = float((err > 2e-3 * scale).float().mean())
        print(f"val {k}: max rel {float(err.max()) / scale:.2e}, frac > 2e-3: {frac_bad:.4f}")
        # sdf_depth switches between crossings on a handful of pixels when the SDF moves by 1e-5; images are smooth
        assert frac_bad <= (0.02 if k == "sdf_depth" else 0.002), (k, frac_bad)
