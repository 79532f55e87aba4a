"""K4 on the tensor cores (csrc/sdf_mlp_tc.cu): the persistent tcgen05 3xTF32 value kernel against the fp32
cuBLAS path (tolerance: 1e-5 + 1e-4 |ref|, as for every fp32 output).  The golden SDF values of the reference
are checked through the same kernel in tests/test_render_gpu.py (sdf_nograd)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0")


def _net(perturb=0.05):
    from gens_b200.config import gens_model_conf
    from gens_b200.implicit_surface import ImplicitSurface
    torch.manual_seed(0)
    surf = ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"]).to(DEV)
    with torch.no_grad():
        for p in surf.sdf_network.parameters():
            p.add_(torch.randn_like(p) * perturb)
    return surf.sdf_network


@pytest.mark.parametrize("n", [1, 127, 128, 129, 148 * 128 + 5, 300_000])
def test_tc_value_kernel_matches_fp32_path(n):
    from gens_b200 import sdf_analytic
    net = _net()
    g = torch.Generator(device=DEV).manual_seed(n)
    dims = [32, 16, 8, 4, 2]
    vols = [torch.randn(1, 4, d, d, d, device=DEV, generator=g) * 0.5 for d in dims]
    pts = torch.rand(n, 3, device=DEV, generator=g) * 2.4 - 1.2   # includes points outside the volumes
    fw = sdf_analytic.FoldedSDF(net)
    try:
        sdf_analytic.USE_TC = False
        ref = sdf_analytic.value_only(net, pts, vols, fw)
        sdf_analytic.USE_TC = True
        out = sdf_analytic.value_only(net, pts, vols, fw)
    finally:
        sdf_analytic.USE_TC = True
    assert out.shape == ref.shape == (n, 1)
    assert not bool(out.isnan().any())
    err = (out - ref).abs()
    assert bool((err <= 1e-5 + 1e-4 * ref.abs()).all()), float(err.max())
    # and the autograd forward of the mirror (what training uses) agrees as well
    with torch.no_grad():
        fwd = net.sdf(pts[:2048], vols)
    assert bool(((out[:2048] - fwd).abs() <= 1e-5 + 1e-4 * fwd.abs()).all())


def test_tc_value_kernel_empty_and_errors():
    from gens_b200 import mlp_tc, sdf_analytic
    net = _net(0.0)
    packed = sdf_analytic.FoldedSDF(net).packed()
    out = mlp_tc.sdf_values(packed, torch.empty(0, 27, device=DEV), torch.empty(0, 100, device=DEV))
    assert out.shape == (0, 1)
    with pytest.raises(RuntimeError):
        mlp_tc.sdf_values(packed, torch.empty(4, 27), torch.empty(4, 100))   # CPU tensors: no fallback
