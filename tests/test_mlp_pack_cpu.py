"""Packing of the SDF network for the tcgen05 kernel (gens_b200/mlp_tc.py), checked on the host: the
k-step stream, replayed by a float64 emulator of the kernel's schedule, must reproduce SDFNetwork.sdf."""
import torch

from gens_b200 import mlp_tc
from gens_b200.config import gens_model_conf
from gens_b200.implicit_surface import ImplicitSurface
from gens_b200.networks import positional_encoding
from gens_b200.sdf_analytic import FoldedSDF
from oracle.torch_oracle import CpuOps


def test_kstep_stream_reproduces_the_network():
    torch.manual_seed(0)
    surf = ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"], ops=CpuOps)
    net = surf.sdf_network
    with torch.no_grad():
        for p in net.parameters():          # leave the geometric init: exercise every weight
            p.add_(torch.randn_like(p) * 0.05)
    packed = mlp_tc.PackedSDF(FoldedSDF(net))
    dims = [16, 8, 4, 2, 2]
    vols = [torch.randn(1, 4, d, d, d) * 0.5 for d in dims]
    pts = torch.rand(300, 3) * 2 - 1
    with torch.no_grad():
        ref = net.sdf(pts, vols)
        fe = positional_encoding(CpuOps.lookup_volume(pts, vols), net.feat_multires)
        pos = positional_encoding(pts * net.scale, net.multires)
    out = mlp_tc.emulate(packed, pos, fe)
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())

    # structure: every block is 16-byte aligned, hi parts are TF32-exact, flags are consistent
    ks = packed.ksteps.cpu()
    assert ks.shape == (packed.n_ksteps, 4) and packed.n_ksteps <= 128
    assert int((ks[:, 3] & 1).ne(0).sum()) == packed.n_layers and int((ks[:, 3] & 2).ne(0).sum()) == packed.n_layers
    assert int((ks[:, 3] & 4).ne(0).sum()) == packed.n_layers - 1
    assert bool(((ks[:, 0] % 16) == 0).all()) and int(ks[:, 1].max()) <= 16384
    off, nbytes = int(ks[0, 0]), int(ks[0, 1])
    hi = packed.wstream[off // 4: off // 4 + nbytes // 8]
    assert bool(((hi.view(torch.int32) & 0x1fff) == 0).all())


def test_reverse_stream_reproduces_gradient_and_hvp():
    """The JVP forward + reverse k-step streams, replayed in float64, against torch.autograd on the same MLP
    written as a function of the encodings: g = d sdf / d(pos, fe), dg = its directional derivative."""
    import math
    torch.manual_seed(1)
    surf = ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"], ops=CpuOps)
    net = surf.sdf_network
    with torch.no_grad():
        for p in net.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    fw = FoldedSDF(net)
    packed, rev = mlp_tc.PackedSDF(fw), mlp_tc.PackedSDFReverse(fw)
    n = 37
    pos = torch.randn(n, 27, dtype=torch.float64) * 0.5
    fe = torch.randn(n, 100, dtype=torch.float64) * 0.5
    dpos, dfe = torch.randn(n, 27, dtype=torch.float64), torch.randn(n, 100, dtype=torch.float64)
    folded = [(w.double(), b.double()) for w, b in net.folded_weights()]

    def mlp(pos_, fe_):
        x = pos_
        last = len(folded) - 1
        for l, (w, b) in enumerate(folded):
            if l in net.skip_in:
                x = torch.cat([x, pos_], -1) / math.sqrt(2)
            if 0 < l:
                x = torch.cat([x, fe_], -1)
            x = torch.nn.functional.linear(x, w, b)
            if l < last:
                x = torch.nn.functional.softplus(x, beta=100)
        return x[:, :1] / net.scale

    pos_r, fe_r = pos.clone().requires_grad_(True), fe.clone().requires_grad_(True)
    y = mlp(pos_r, fe_r)
    gp, gf = torch.autograd.grad(y.sum(), (pos_r, fe_r), create_graph=True)
    dot = (gp * dpos).sum() + (gf * dfe).sum()
    dgp, dgf = torch.autograd.grad(dot, (pos_r, fe_r))

    sdf, g_pos, g_fe = mlp_tc.emulate_grad(packed, rev, torch.cat([pos, dpos]).float(), torch.cat([fe, dfe]).float(), n)

    def close(a, b, what):
        err = (a.double() - b).abs().max().item()
        assert err <= 2e-5 * max(1.0, b.abs().max().item()), (what, err)

    close(sdf, y.detach(), "sdf")
    close(g_pos[:n], gp.detach(), "g_pos")
    close(g_fe[:n], gf.detach(), "g_fe")
    close(g_pos[n:], dgp, "dg_pos")
    close(g_fe[n:], dgf, "dg_fe")
