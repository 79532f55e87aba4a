"""K9 (csrc/tv_reg.cu): the one-pass masked total variation against the ATen-op formulation of
ImplicitSurface.tv_regularization (reference implicit_surface.py:135-150), which the golden render fixtures
pin.  Tolerance: 1e-6 + 1e-4 |ref| (the kernel accumulates in fp64, the reference in fp32)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _surf():
    from gens_b200.config import gens_model_conf
    from gens_b200.implicit_surface import ImplicitSurface
    torch.manual_seed(0)
    return ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"]).to(DEV)


def _torch_tv(surf, vols, masks):
    """The reference formulation: forced by asking for a graph."""
    vs = [v.clone().requires_grad_(True) for v in vols]
    with torch.enable_grad():
        return surf.tv_regularization(vs, masks).detach()


@pytest.mark.parametrize("dims", [[32, 16, 8, 4, 2], [64, 33, 7, 1], [128]])
@pytest.mark.parametrize("with_mask", [True, False])
def test_tv_kernel_matches_aten_formulation(cuda_lib, dims, with_mask):
    from gens_b200 import projector
    surf = _surf()
    g = torch.Generator(device=DEV).manual_seed(sum(dims))
    vols = [torch.randn(1, 4, d, d, d, device=DEV, generator=g) for d in dims]
    masks = [(torch.rand(1, 1, d, d, d, device=DEV, generator=g) > 0.4).float() for d in dims] if with_mask else None
    projector.clear_caches()
    with torch.no_grad():
        got = surf.tv_regularization(vols, masks)
    ref = _torch_tv(surf, vols, masks)
    assert got.shape == ref.shape == () and got.dtype == torch.float32
    assert abs(float(got) - float(ref)) <= 1e-6 + 1e-4 * abs(float(ref)), (float(got), float(ref))


def test_tv_edge_cases_and_cache(cuda_lib):
    from gens_b200 import _lib, projector
    surf = _surf()
    d = 16
    vol = torch.randn(1, 4, d, d, d, device=DEV)
    empty = torch.zeros(1, 1, d, d, d, device=DEV)
    projector.clear_caches()
    with torch.no_grad():
        assert float(surf.tv_regularization([vol], [empty])) == 0.0     # no valid pair: 0 / 1e-8
        a = surf.tv_regularization([vol])
        n0 = _lib.LAUNCHES
        b = surf.tv_regularization([vol])                                 # same tensors, same version: cached
        assert _lib.LAUNCHES == n0 and float(a) == float(b)
        vol.mul_(2.0)                                                     # in-place update bumps _version
        c = surf.tv_regularization([vol])
        assert _lib.LAUNCHES == n0 + 1
    assert abs(float(c) - 2.0 * float(a)) <= 1e-5 * float(c)
    # a volume that needs a gradient (fine-tuning) keeps the differentiable ATen formulation
    vg = vol.clone().requires_grad_(True)
    t = surf.tv_regularization([vg])
    assert t.requires_grad
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            surf.tv_regularization([vol.cpu()])                           # no CPU fallback
