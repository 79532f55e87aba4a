"""Generate golden input/output vectors by running the UNMODIFIED reference on CPU.

Runs only in the build container (needs /root/reference); the .npz fixtures it writes are
committed so that tests on the GPU box never touch the reference tree.

    python tests/golden/make_golden.py [volume] [render] ...

Fixtures
  volume_agg.npz   Volume.agg_mean_var (reference models/modules/volume.py:13-63) on a
                   96x128, 3-view scene, volume_dims [32,16,8,4,2]; also per-view validity
                   masks (single-view runs with min_vis_view=0) and the normalised grid the
                   reference handed to F.grid_sample (captured, to derive corner indices).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from gens_b200.synthetic import make_scene  # noqa: E402


class Conf(dict):
    """Minimal stand-in for the pyhocon ConfigTree the reference reads (pyhocon is absent)."""

    def _get(self, key):
        node = self
        for part in key.split("."):
            node = node[part]
        return node

    def get_list(self, key):
        return list(self._get(key))

    def get_int(self, key):
        return int(self._get(key))

    def get_float(self, key):
        return float(self._get(key))

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        return Conf(v) if isinstance(v, dict) and not isinstance(v, Conf) else v


def load_ref_module(rel_path: str, name: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel_path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_volume():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    volume_mod = load_ref_module("models/modules/volume.py", "ref_volume")
    dims = [32, 16, 8, 4, 2]
    vol = volume_mod.Volume(Conf(volume_dims=dims))
    scene = make_scene(96, 128, 3, seed=0)

    captured = []
    real_grid_sample = F.grid_sample

    def spy(inp, grid, *a, **k):
        captured.append(grid.detach().clone())
        return real_grid_sample(inp, grid, *a, **k)

    volume_mod.F.grid_sample = spy
    try:
        volumes, masks = vol.agg_mean_var(scene.features, scene.intrs, scene.c2ws)
    finally:
        volume_mod.F.grid_sample = real_grid_sample
    out = {
        "dims": np.array(dims), "hw": np.array(scene.hw),
        "intrs": scene.intrs.numpy(), "c2ws": scene.c2ws.numpy(),
    }
    for i, d in enumerate(dims):
        out[f"feat{i}"] = scene.features[i].numpy()
        out[f"volume{i}"] = volumes[i][0].numpy()
        out[f"mask{i}"] = masks[i][0, 0].numpy()
        out[f"grid{i}"] = captured[i][:, 0].numpy()  # (nv, D^3, 2) normalised x,y
        # exact visible-view count from the strict threshold at 0,1,2
        cnt = np.zeros((d, d, d), np.int32)
        for t in range(3):
            _, mk = vol.agg_mean_var(scene.features, scene.intrs, scene.c2ws, min_vis_view=t)
            cnt += mk[i][0, 0].numpy().astype(np.int32)
        out[f"count{i}"] = cnt
        pv = []
        for v in range(3):
            _, mk = vol.agg_mean_var([f[v:v + 1] for f in scene.features], scene.intrs[v:v + 1],
                                     scene.c2ws[v:v + 1], min_vis_view=0)
            pv.append(mk[i][0, 0].numpy().astype(np.uint8))
        out[f"viewmask{i}"] = np.stack(pv)
        assert (out[f"viewmask{i}"].sum(0) == cnt).all(), "per-view masks inconsistent with counts"
    np.savez_compressed(os.path.join(HERE, "volume_agg.npz"), **out)
    print("volume_agg.npz written; mask fill per scale:",
          [float(out[f"mask{i}"].mean()) for i in range(len(dims))])


RENDER_DIMS = [32, 16, 8, 4, 2]
RENDER_HW = (96, 128)


def render_inputs():
    """Deterministic inputs of the render-half fixtures; tests rebuild them from the same seeds."""
    from gens_b200.synthetic import make_reg_volumes
    scene = make_scene(RENDER_HW[0], RENDER_HW[1], 3, seed=11)
    volumes = make_reg_volumes(RENDER_DIMS, seed=11)
    return scene, volumes


def golden_render():
    sys.path.insert(0, HERE)
    import ref_shims
    ref_shims.install()
    import models.modules.implicit_surface as IS
    import models.modules.projector as PJ
    volume_mod = load_ref_module("models/modules/volume.py", "ref_volume")

    torch.manual_seed(0)
    surf = IS.ImplicitSurface(Conf(ref_shims.REF_CONF))
    # random-init colour net is fine; nudge the variance so inv_s is the init value exp(3)
    scene, volumes = render_inputs()
    _, masks = volume_mod.Volume(Conf(volume_dims=RENDER_DIMS)).agg_mean_var(scene.features, scene.intrs, scene.c2ws)
    out = {"mask_fill": np.array([float(m.mean()) for m in masks])}
    for k, v in surf.state_dict().items():
        out["sd/" + k] = v.numpy()
    for i, m in enumerate(masks):
        out[f"mask{i}"] = m[0, 0].numpy().astype(np.uint8)

    rays_o, rays_d = scene.rays(step=8)
    sel = torch.arange(0, rays_o.shape[0], 4)[:48]
    rays_o, rays_d = rays_o[sel].contiguous(), rays_d[sel].contiguous()
    out["rays_o"], out["rays_d"] = rays_o.numpy(), rays_d.numpy()

    # ---- lookup_volume: nearest masks / trilinear features, points inside and outside the cube
    g = torch.Generator().manual_seed(5)
    pts = (torch.rand(4000, 3, generator=g) * 2.6 - 1.3)
    pts[:8] = torch.tensor([[-1., -1, -1], [1, 1, 1], [0, 0, 0], [1, -1, 0.5], [-1.0001, 0, 0], [0.999999, 0.5, -0.5],
                            [0.03125, 0.0625, -0.09375], [1.2, 1.2, 1.2]])
    out["lv_pts"] = pts.numpy()
    out["lv_nearest"] = PJ.lookup_volume(pts, masks, sample_mode="nearest").numpy()
    out["lv_feat"] = PJ.lookup_volume(pts, volumes, sample_mode="grad").detach().numpy()

    # ---- SDF network: value, gradient, second-order "smooth" on points along the rays
    z = scene.near + (scene.far - scene.near) * torch.linspace(0, 1, 40)[None, :]
    spts = (rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]).reshape(-1, 3)
    out["sdf_pts"] = spts.numpy()
    out["sdf_out"] = surf.sdf_network(spts, volumes).detach().numpy()
    gr, sm = surf.sdf_network.gradient(spts.clone(), volumes)
    out["sdf_grad"], out["sdf_smooth"] = gr.detach().numpy(), sm.detach().numpy()

    # ---- lookup_feature + colour network
    fv, rd, mk = PJ.lookup_feature(spts, scene.imgs, scene.intrs, scene.c2ws, scene.features)
    out["lf_feat"], out["lf_raydiff"], out["lf_mask"] = fv.detach().numpy(), rd.numpy(), mk.numpy()
    out["color"] = surf.color_network(fv, rd, mk).detach().numpy()

    # ---- full render (perturb = 0); the 1024 random "sparse" points come from the global RNG
    captured = {}
    real_core = surf.render_core

    def spy(rays_o_, rays_d_, z_vals, *a, **k):
        captured["z_vals"] = z_vals.detach().clone()
        return real_core(rays_o_, rays_d_, z_vals, *a, **k)

    surf.render_core = spy
    torch.manual_seed(123)
    res = surf.render(rays_o, rays_d, scene.near, scene.far, volumes, masks, scene.imgs, scene.features,
                      scene.features, scene.intrs, scene.c2ws, 1.0, None)
    out["z_vals"] = captured["z_vals"].numpy()
    for k, v in res.items():
        out["render/" + k] = v.detach().numpy()
    # coarse stage pieces for unit tests of up_sample / sample_pdf
    with torch.no_grad():
        z64 = scene.near + (scene.far - scene.near) * torch.linspace(0, 1, 64)[None, :]
        z64 = z64.repeat(rays_o.shape[0], 1)
        p64 = (rays_o[:, None, :] + rays_d[:, None, :] * z64[..., None]).reshape(-1, 3)
        m64 = PJ.lookup_volume(p64, masks, sample_mode="nearest").any(dim=-1)
        sdf64 = torch.ones(p64.shape[0], 1) * 100
        sdf64[m64] = surf.sdf_network.sdf(p64[m64], volumes)
        sdf64 = sdf64.reshape(-1, 64)
        new_z = surf.up_sample(rays_o, rays_d, z64, sdf64, 16, masks, 64)
    out["up_z64"], out["up_sdf64"], out["up_new_z"] = z64.numpy(), sdf64.numpy(), new_z.numpy()
    np.savez_compressed(os.path.join(HERE, "render.npz"), **out)
    print("render.npz written; mask fill", out["mask_fill"], "valid rays", int(res["valid_mask"].sum()),
          "of", rays_o.shape[0], "weight_sum mean", float(res["weight_sum"].mean()))


BIG_DIMS = [128, 64, 32, 16, 8]
BIG_HW = (240, 320)


def big_render_inputs():
    from gens_b200.synthetic import make_reg_volumes
    scene = make_scene(BIG_HW[0], BIG_HW[1], 3, seed=13)
    volumes = make_reg_volumes(BIG_DIMS, seed=13)
    rays_o, rays_d = scene.rays(step=16)
    sel = torch.arange(0, rays_o.shape[0], 6)[:48]
    return scene, volumes, rays_o[sel].contiguous(), rays_d[sel].contiguous()


def golden_render_big():
    """render() of the unmodified reference through a pyramid whose finest scale is 128^3 (8.4 M voxels per channel:
    the look-up kernels' large-volume indexing, the 5-scale mask pyramid built at 240x320), 48 rays.  Only outputs and
    the bit-packed masks are stored; inputs come back from seeds, the weights from render.npz's state_dict."""
    sys.path.insert(0, HERE)
    import ref_shims
    ref_shims.install()
    import models.modules.implicit_surface as IS
    volume_mod = load_ref_module("models/modules/volume.py", "ref_volume")
    small = np.load(os.path.join(HERE, "render.npz"))
    torch.manual_seed(0)
    surf = IS.ImplicitSurface(Conf(ref_shims.REF_CONF))
    surf.load_state_dict({k[3:]: torch.from_numpy(small[k]) for k in small.files if k.startswith("sd/")}, strict=True)
    scene, volumes, rays_o, rays_d = big_render_inputs()
    with torch.no_grad():
        _, masks = volume_mod.Volume(Conf(volume_dims=BIG_DIMS)).agg_mean_var(scene.features, scene.intrs, scene.c2ws)
    out = {"mask_fill": np.array([float(m.mean()) for m in masks])}
    # the reference's own torch.inverse(c2ws) on this CPU (volume.py:34): LAPACK results can differ in the last bit
    # between hosts, and a last-bit difference flips borderline voxels of a 128^3 mask
    out["w2c"] = torch.inverse(scene.c2ws).numpy()
    for i, m in enumerate(masks):
        out[f"maskbits{i}"] = np.packbits(m[0, 0].numpy().astype(np.uint8).reshape(-1))
    captured = {}
    real_core = surf.render_core

    def spy(rays_o_, rays_d_, z_vals, *a, **k):
        captured["z_vals"] = z_vals.detach().clone()
        return real_core(rays_o_, rays_d_, z_vals, *a, **k)

    surf.render_core = spy
    torch.manual_seed(123)
    res = surf.render(rays_o, rays_d, scene.near, scene.far, volumes, masks, scene.imgs, scene.features,
                      scene.features, scene.intrs, scene.c2ws, 1.0, None)
    out["z_vals"] = captured["z_vals"].numpy()
    for k, v in res.items():
        out["render/" + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "render_big.npz"), **out)
    print("render_big.npz written; mask fill", out["mask_fill"], "valid rays", int(res["valid_mask"].sum()), "of",
          rays_o.shape[0], "weight_sum mean", float(res["weight_sum"].mean()))


TRAIN_DIMS = [32, 16, 8, 8, 4]   # every scale keeps visible voxels (an all-masked scale gives the reference a NaN TV gradient)
TRAIN_HW = (96, 128)
TRAIN_NV = 5          # config 3: 4 source views (confs/gens.conf:9)
TRAIN_RAYS = 32
# train.loss block of confs/gens.conf:47-58
LOSS_CONF = dict(color_weight=1.0, sparse_scale_factor=100.0, sparse_weight=0.02, igr_weight=0.1, mfc_weight=1.0,
                 smooth_weight=0.0001, tv_weight=0.0001, depth_weight=0.0, pseudo_sdf_weight=1.0,
                 pseudo_depth_weight=0.05)


def train_inputs():
    """Deterministic inputs of the training fixture (config-3 shape, mini scene); tests rebuild them from the seeds."""
    from gens_b200.synthetic import make_reg_volumes
    scene = make_scene(TRAIN_HW[0], TRAIN_HW[1], TRAIN_NV, seed=31)
    volumes = make_reg_volumes(TRAIN_DIMS, seed=31)
    g = torch.Generator().manual_seed(77)
    ro, rd = scene.rays(step=1)
    sel = torch.randperm(ro.shape[0], generator=g)[:TRAIN_RAYS]
    rays_o, rays_d = ro[sel].contiguous(), rd[sel].contiguous()
    pseudo = torch.rand(256, 3, generator=g) * 1.0 - 0.5
    target = torch.rand(TRAIN_RAYS, 3, generator=g)
    return scene, volumes, rays_o, rays_d, pseudo, target


def golden_train():
    """The reference's training step on the mini scene: ImplicitSurface.forward("train") (implicit_surface.py:472-499,
    incl. pseudo_pts) -> Loss.forward (models/losses/loss.py:23-84, weights of confs/gens.conf) -> backward().
    Records every output, the loss terms and the gradients w.r.t. the five feature maps, the five volumes and every
    MLP parameter -- the second-order path of SDFNetwork.gradient (sdf_network.py:131-153) included."""
    sys.path.insert(0, HERE)
    import ref_shims
    ref_shims.install()
    import models.modules.implicit_surface as IS
    import models.losses.loss as LS
    volume_mod = load_ref_module("models/modules/volume.py", "ref_volume")

    torch.manual_seed(0)
    surf = IS.ImplicitSurface(Conf(ref_shims.REF_CONF))
    from gens_b200.config import Conf as FullConf  # the ConfigTree look-alike with `default=` getters
    loss_fn = LS.Loss(FullConf(LOSS_CONF))
    # geometric initialisation zeroes the weights that read the volume features; a small perturbation makes the
    # gradients w.r.t. the volumes a real signal instead of second-order crumbs (the state_dict is recorded)
    gp = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for n_, p_ in surf.sdf_network.named_parameters():
            if n_.endswith("weight_v"):
                p_.add_(0.004 * torch.randn(p_.shape, generator=gp))
    scene, volumes, rays_o, rays_d, pseudo, target = train_inputs()
    with torch.no_grad():
        _, masks = volume_mod.Volume(Conf(volume_dims=TRAIN_DIMS)).agg_mean_var(scene.features, scene.intrs, scene.c2ws)
    vols = [v.clone().requires_grad_(True) for v in volumes]
    feats = [f.clone().requires_grad_(True) for f in scene.features]
    ipts = {"imgs": scene.imgs, "intrs": scene.intrs, "c2ws": scene.c2ws, "rays_o": rays_o, "rays_d": rays_d,
            "near": scene.near, "far": scene.far, "pseudo_pts": pseudo}
    captured = {}
    real_core = surf.render_core

    def spy(rays_o_, rays_d_, z_vals, *a, **k):
        captured["z_vals"] = z_vals.detach().clone()
        return real_core(rays_o_, rays_d_, z_vals, *a, **k)

    surf.render_core = spy
    torch.manual_seed(123)
    res = surf("train", ipts, vols, masks, feats, feats, cos_anneal_ratio=0.7, step=3)
    losses = loss_fn(res, {"color": target}, step=3)
    losses["loss"].backward()

    out = {"mask_fill": np.array([float(m.mean()) for m in masks])}
    for k, v in surf.state_dict().items():
        out["sd/" + k] = v.numpy()
    for i, m in enumerate(masks):
        out[f"mask{i}"] = m[0, 0].numpy().astype(np.uint8)
    out["z_vals"] = captured["z_vals"].numpy()  # the 128 depths per ray the reference composited (after up-sampling)
    for k, v in res.items():
        out["out/" + k] = v.detach().numpy()
    for k, v in losses.items():
        out["loss/" + k] = np.asarray(v.detach().numpy())
    for n, p_ in surf.named_parameters():
        if p_.grad is not None:
            out["grad/param/" + n] = p_.grad.numpy()
    for i, v in enumerate(vols):
        out[f"grad/volume{i}"] = v.grad.numpy()
    for i, f in enumerate(feats):
        out[f"grad/feature{i}"] = (f.grad if f.grad is not None else torch.zeros_like(f)).numpy()
    np.savez_compressed(os.path.join(HERE, "train.npz"), **out)
    print("train.npz written; loss terms", {k: float(v) for k, v in losses.items()}, "mask fill", out["mask_fill"],
          "valid rays", int(res["valid_mask"].sum()), "of", TRAIN_RAYS,
          "| grad norms: params", float(sum((p_.grad ** 2).sum() for p_ in surf.parameters() if p_.grad is not None)) ** 0.5,
          "volumes", [float(v.grad.norm()) for v in vols], "features", [float(f.grad.norm()) for f in feats if f.grad is not None])


def golden_reg_network():
    """reg_network.npz: outputs of the UNMODIFIED reference RegNetwork (models/modules/reg_network.py:105-166) on the
    seeded volumes / parameters of tests/test_reg_network_cpu.py (its seeded_state / seeded_volumes rules, which do not
    depend on construction order): strided samples of every output + its sum."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from test_reg_network_cpu import DIMS, seeded_state, seeded_volumes
    mod = load_ref_module("models/modules/reg_network.py", "_golden_reg_network")
    net = mod.RegNetwork(Conf({"d_voluem": [8] * 5, "d_out": [4] * 5, "d_base": 8}))
    net.load_state_dict(seeded_state(net))
    with torch.no_grad():
        outs = net(seeded_volumes(DIMS))
    stride = [4, 2, 1, 1, 1]
    rec = {"stride": np.array(stride), "sum": np.array([float(o.double().sum()) for o in outs])}
    for i, o in enumerate(outs):
        rec[f"out{i}"] = o[0, :, ::stride[i], ::stride[i], ::stride[i]].numpy()
    np.savez_compressed(os.path.join(HERE, "reg_network.npz"), **rec)
    print("reg_network.npz:", {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["volume", "render"]
    if "reg_network" in which:
        golden_reg_network()
    if "train" in which:
        golden_train()
    if "render_big" in which:
        golden_render_big()
    if "volume" in which:
        golden_volume()
    if "render" in which:
        golden_render()
