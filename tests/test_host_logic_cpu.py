"""The product's HOST logic (gens_b200.implicit_surface / networks: sampling schedule, NeuS weights,
compositing, zero crossing, output dictionary) checked on CPU against the reference's golden render,
with the CUDA look-ups swapped for the ATen-on-CPU provider of oracle/.  No GPU, no CUDA extension."""
import numpy as np
import torch

from gens_b200.config import gens_model_conf
from gens_b200.implicit_surface import ImplicitSurface, sample_pdf
from gens_b200.synthetic import make_reg_volumes, make_scene
from oracle.torch_oracle import CpuOps
from parity import JUMPY, mismatch

DIMS = [32, 16, 8, 4, 2]


def _setup(golden_dir):
    g = np.load(f"{golden_dir}/render.npz")
    surf = ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"], ops=CpuOps)
    surf.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd/")}, strict=True)
    scene = make_scene(96, 128, 3, seed=11)
    volumes = make_reg_volumes(DIMS, seed=11)
    masks = [torch.from_numpy(g[f"mask{i}"].astype(np.float32))[None, None] for i in range(5)]
    return g, surf, scene, volumes, masks


def test_render_host_logic_matches_reference_golden(golden_dir):
    g, surf, scene, volumes, masks = _setup(golden_dir)
    ro, rd = torch.from_numpy(g["rays_o"])[:12], torch.from_numpy(g["rays_d"])[:12]
    # the reference rendered 48 rays in one call; per-ray outputs do not depend on the batch
    torch.manual_seed(123)
    res = surf.render(ro, rd, scene.near, scene.far, volumes, masks, scene.imgs, scene.features, scene.features,
                      scene.intrs, scene.c2ws, 1.0, None)
    assert sorted(res.keys()) == sorted(k[7:] for k in g.files if k.startswith("render/"))
    for k in ("color_fine", "render_depth", "weights", "weight_sum", "sdf_depth", "normal", "gradients",
              "inside_sphere", "valid_mask", "mid_inside_sphere", "ref_gray_val", "sampled_gray_val"):
        ref = g["render/" + k]
        ref = ref[:, :12] if k in ("ref_gray_val", "sampled_gray_val") else ref[:12]
        msg = mismatch(k, res[k], ref, atol_scale=1e-5, outlier_frac=JUMPY.get(k, 0.0))
        assert msg is None, msg


def test_up_sample_and_sdf_nograd(golden_dir):
    g, surf, scene, volumes, masks = _setup(golden_dir)
    ro, rd = torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])
    with torch.no_grad():
        new_z = surf.up_sample(ro, rd, torch.from_numpy(g["up_z64"]), torch.from_numpy(g["up_sdf64"]), 16, masks, 64)
    assert np.allclose(new_z.numpy(), g["up_new_z"], rtol=1e-4, atol=1e-5)
    pts = torch.from_numpy(g["sdf_pts"])
    assert np.allclose(surf.sdf_network.sdf_nograd(pts, volumes).numpy(), g["sdf_out"][:, :1], rtol=1e-4, atol=1e-6)
    grad, smooth = surf.sdf_network.gradient(pts.clone(), volumes)
    assert np.allclose(grad.detach().numpy(), g["sdf_grad"], rtol=1e-4, atol=1e-6)
    assert np.allclose(smooth.detach().numpy(), g["sdf_smooth"], rtol=1e-4, atol=1e-4 * np.abs(g["sdf_smooth"]).max())


def test_sample_pdf_edge_cases():
    bins = torch.linspace(0, 1, 9)[None].repeat(3, 1)
    w = torch.zeros(3, 8)
    w[1, 3] = 1.0           # all mass in one bin
    w[2] = 1.0              # uniform
    z = sample_pdf(bins, w, 16, det=True)
    assert z.shape == (3, 16) and torch.all(z[:, 1:] >= z[:, :-1])
    assert torch.all((z[1] > bins[1, 3] - 1e-4) & (z[1] < bins[1, 4] + 1e-4))
    assert torch.allclose(z[2], torch.linspace(0.5 / 16, 1 - 0.5 / 16, 16), atol=1e-5)
