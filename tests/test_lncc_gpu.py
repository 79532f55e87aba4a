"""K11 (csrc/lncc.cu) -- compute_LNCC drop-in (reference models/losses/ncc.py:7-50) through the C ABI: forward and
both patch gradients against the oracle restatement (oracle/torch_oracle.compute_lncc, itself pinned to the
reference's own loss terms by tests/test_oracle_golden.py) and against the patches / mfc_loss the reference recorded
for the training fixture."""
import numpy as np
import pytest
import torch

from gens_b200.losses import compute_LNCC
from oracle import torch_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("ns,b,patch,c", [(4, 64, 11, 12), (2, 7, 11, 12), (3, 33, 5, 4)])
def test_lncc_forward_backward_match_oracle(cuda_lib, ns, b, patch, c):
    g = torch.Generator().manual_seed(ns * 100 + b)
    p = patch * patch
    ref = torch.rand(1, b, p, c, generator=g)
    # correlated sources (so that cc is neither 0 nor saturated), one view anti-correlated, one constant patch
    src = 0.6 * ref + 0.4 * torch.rand(ns, b, p, c, generator=g)
    src[0, : b // 3] = 1.0 - src[0, : b // 3]
    src[-1, 0] = 0.25
    w = torch.rand(b, 1, generator=g)

    rc, sc = ref.clone().requires_grad_(True), src.clone().requires_grad_(True)
    want = torch_oracle.compute_lncc(rc, sc)
    (want * w).sum().backward()

    rg, sg = ref.to(DEV).requires_grad_(True), src.to(DEV).requires_grad_(True)
    got = compute_LNCC(rg, sg)
    (got * w.to(DEV)).sum().backward()
    assert got.shape == (b, 1)
    # sums of 121 products in a different order than the CPU reduction: 1e-5 absolute on a [0, 2] score
    assert torch.allclose(got.cpu(), want.detach(), rtol=1e-4, atol=1e-5), float((got.cpu() - want).abs().max())
    # ray 0 holds the constant source patch: its variance is 0 in exact arithmetic and fp32 cancellation noise of
    # either sign in any implementation (cc = noise^2 / (noise + 1e-5)): the score is pinned above, the gradient of
    # that ray is only required to be finite -- the reference's own autograd returns noise there as well
    for name, a, e in (("ref", rg.grad.cpu()[:, 1:], rc.grad[:, 1:]), ("src", sg.grad.cpu()[:, 1:], sc.grad[:, 1:])):
        scale = float(e.abs().max())
        assert float((a - e).abs().max()) <= 2e-4 * scale + 1e-7, (name, float((a - e).abs().max()), scale)
    assert bool(torch.isfinite(rg.grad).all()) and bool(torch.isfinite(sg.grad).all())


def test_lncc_matches_reference_recorded_patches(cuda_lib, golden_dir):
    """On the ref_gray_val / sampled_gray_val the unmodified reference produced for the training fixture, the kernel's
    scores give the mfc_loss the reference's own Loss computed (tests/golden/train.npz)."""
    g = np.load(f"{golden_dir}/train.npz")
    ref = torch.from_numpy(g["out/ref_gray_val"]).to(DEV)
    src = torch.from_numpy(g["out/sampled_gray_val"]).to(DEV)
    ncc = compute_LNCC(ref, src)
    mask = torch.from_numpy(g["out/valid_mask"]).float().to(DEV) * torch.from_numpy(g["out/mid_inside_sphere"]).to(DEV)
    mfc = 0.5 * ((ncc * mask).sum(dim=0) / (mask.sum(dim=0) + 1e-8)).squeeze(-1)
    assert abs(float(mfc) - float(g["loss/mfc_loss"])) <= 1e-5, (float(mfc), float(g["loss/mfc_loss"]))
    want = torch_oracle.compute_lncc(ref.cpu(), src.cpu())
    assert torch.allclose(ncc.cpu(), want, rtol=1e-4, atol=1e-5)
