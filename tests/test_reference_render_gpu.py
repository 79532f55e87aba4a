"""ImplicitSurface.render through the CONFIG-2 volume pyramid (256^3 ... 16^3, 480x640, 3 views) against the UNMODIFIED
reference on the same GPU (baseline/_ref, its own gridsample_grad2 extension): the large-volume indexing of the
look-up kernels, the mask pyramid K1 builds at BASELINE sizes and the whole fused inference march, per-ray outputs
compared ray by ray.  (The golden fixtures stop at a 128^3 pyramid; VERDICT r01 "weak" item 1.)

Skipped when baseline/_ref is not staged (python baseline/setup_ref.py)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_runtime  # noqa: E402

from gens_b200.config import gens_model_conf  # noqa: E402
from gens_b200.implicit_surface import ImplicitSurface  # noqa: E402
from gens_b200.synthetic import make_scene  # noqa: E402
from gens_b200.volume import Volume  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_runtime.available(), reason=ref_runtime.why_unavailable() or "-")]
DIMS = [256, 128, 64, 32, 16]


def _smooth_volumes(dev):
    g = torch.Generator(device=dev).manual_seed(1)
    vols = []
    for d in DIMS:
        lo = max(d // 8, 2)
        base = torch.randn(1, 4, lo, lo, lo, device=dev, generator=g) * 0.5
        vols.append(torch.nn.functional.interpolate(base, size=(d, d, d), mode="trilinear", align_corners=True).contiguous())
    return vols


def test_render_through_config2_volumes_matches_reference(cuda_lib):
    dev = torch.device("cuda:0")
    sc = make_scene(480, 640, 3, seed=0).to(dev)
    vols = _smooth_volumes(dev)
    with torch.no_grad():
        _, masks = Volume(volume_dims=DIMS).agg_mean_var(sc.features, sc.intrs, sc.c2ws)
    torch.manual_seed(0)
    ours = ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"]).to(dev).eval()
    ns = ref_runtime.load()  # fresh, un-patched import of the staged reference
    try:
        assert ns.implicit_surface.lookup_volume.__module__ == "models.modules.projector"
        ref = ns.implicit_surface.ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"]).to(dev).eval()
        ref.load_state_dict(ours.state_dict())
        ro, rd = sc.rays(step=1)
        sel = torch.arange(0, ro.shape[0], ro.shape[0] // 1024, device=ro.device)[:1024]
        ro, rd = ro[sel].to(dev).contiguous(), rd[sel].to(dev).contiguous()
        outs = []
        for model in (ref, ours):
            torch.manual_seed(7)
            res = []
            with torch.no_grad():
                for a in range(0, ro.shape[0], 256):
                    r = model.render(ro[a:a + 256], rd[a:a + 256], sc.near, sc.far, vols, masks, sc.imgs, sc.features,
                                     sc.features, sc.intrs, sc.c2ws, 1.0, None)
                    res.append({k: r[k].detach().float() for k in ("color_fine", "render_depth", "normal", "weight_sum",
                                                                    "sdf_depth", "valid_mask", "weights")})
            outs.append({k: torch.cat([x[k] for x in res]) for k in res[0]})
    finally:
        ref_runtime.purge()
    o_ref, o_our = outs
    assert torch.equal(o_our["valid_mask"] > 0, o_ref["valid_mask"] > 0)
    assert float(o_ref["weight_sum"].mean()) > 0.2, "the test scene must have a surface"
    for k, tight, share_lim in (("color_fine", 2e-3, 0.01), ("render_depth", 1e-3, 0.01), ("normal", 2e-3, 0.01),
                                ("weight_sum", 1e-3, 0.01), ("weights", 2e-3, 0.01), ("sdf_depth", 2e-3, 0.02)):
        a, b = o_our[k], o_ref[k]
        err = (a - b).abs() / b.abs().max().clamp_min(1e-12)
        share = float((err > tight).float().mean())
        print(f"{k}: max {float(err.max()):.2e} of scale, share beyond {tight:g}: {share:.4f}")
        # the shipped 3xTF32 importance sampling places samples <= 1e-4 away from the reference's: per-ray outputs within
        # `tight` of the tensor's scale except for the stated share of rays whose sample flipped across a voxel face
        assert share <= share_lim, (k, share)
