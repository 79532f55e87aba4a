"""K13 (csrc/conv3d.cu) and the CUDA path of the volume regulariser (gens_b200/reg_network.py)."""
import copy
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _k13(x, w, bias=None, lo=None, hi=None, stats=True):
    from gens_b200 import _lib
    c_out, c_in = w.shape[:2]
    _, _, d, h, wd = x.shape
    pk = w.permute(1, 3, 4, 2, 0).contiguous()
    y = torch.full((1, c_out, d, h, wd), float("nan"), device=DEV)
    st = torch.zeros(2 * c_out, device=DEV, dtype=torch.float64)
    null = ctypes.c_void_p(0)
    _lib.check(_lib.lib().gens_conv3d_k3(
        _lib.ptr(x), _lib.ptr(lo) if lo is not None else null, _lib.ptr(hi) if hi is not None else null, _lib.ptr(pk),
        _lib.ptr(bias) if bias is not None else null, c_in, c_out, d, h, wd, _lib.ptr(y),
        _lib.ptr(st) if stats else null, _lib.stream_ptr(DEV)), "gens_conv3d_k3")
    torch.cuda.synchronize()
    return y, st


@pytest.mark.parametrize("c_in,c_out,shape,bias", [
    (8, 8, (16, 24, 64), False), (16, 16, (8, 16, 32), False), (8, 4, (12, 16, 96), True), (32, 4, (4, 8, 32), True),
    (8, 8, (6, 12, 40), False), (8, 16, (5, 9, 33), True), (16, 8, (1, 1, 1), False)])
def test_k13_conv_matches_float64_convolution(cuda_lib, c_in, c_out, shape, bias):
    """Whole-volume case incl. partial tiles (sizes that are no multiple of the 4 x 8 x 32 tile), bias, moments."""
    g = torch.Generator().manual_seed(c_in * 100 + c_out)
    x = torch.randn(1, c_in, *shape, generator=g).to(DEV)
    w = (torch.randn(c_out, c_in, 3, 3, 3, generator=g) / (27 * c_in) ** 0.5).to(DEV)
    b = torch.randn(c_out, generator=g).to(DEV) if bias else None
    want = F.conv3d(x.double(), w.double(), b.double() if bias else None, padding=1)
    got, st = _k13(x, w, b)
    assert not torch.isnan(got).any()
    assert torch.all((got.double() - want).abs() <= 1e-6 * want.abs().max() + 1e-5 * want.abs())
    sums = torch.cat([want.sum(dim=(0, 2, 3, 4)), (want * want).sum(dim=(0, 2, 3, 4))])
    assert torch.all((st - sums).abs() <= 1e-5 * sums.abs() + 1e-4 * want[0, 0].numel() ** 0.5)
    # without the moment buffer
    got2, _ = _k13(x, w, b, stats=False)
    assert torch.equal(got, got2)


def test_k13_conv_on_a_slab_reads_its_halo_planes(cuda_lib):
    g = torch.Generator().manual_seed(5)
    full = torch.randn(1, 8, 24, 16, 64, generator=g).to(DEV)
    w = (torch.randn(8, 8, 3, 3, 3, generator=g) / 15).to(DEV)
    want = F.conv3d(full.double(), w.double(), None, padding=1)
    for a0, a1 in ((0, 8), (8, 16), (16, 24), (4, 7)):
        lo = full[0, :, a0 - 1].contiguous() if a0 > 0 else None
        hi = full[0, :, a1].contiguous() if a1 < 24 else None
        got, _ = _k13(full[:, :, a0:a1].contiguous(), w, None, lo, hi)
        ref = want[:, :, a0:a1]
        assert torch.all((got.double() - ref).abs() <= 1e-6 * ref.abs().max() + 1e-5 * ref.abs()), (a0, a1)
    # an interior slab WITHOUT its halo planes must differ (the planes are really read)
    got, _ = _k13(full[:, :, 8:16].contiguous(), w)
    assert not torch.allclose(got.double()[:, :, 0], want[:, :, 8], atol=1e-4)


def _strided(x, w, transposed, halo=None):
    from gens_b200 import _lib
    _, c_in, d, h, wd = x.shape
    if transposed:
        c_out, pk, shape = w.shape[1], w.permute(0, 2, 3, 4, 1).contiguous(), (2 * d, 2 * h, 2 * wd)
        fn = _lib.lib().gens_deconv3d_k3s2
    else:
        c_out, pk, shape = w.shape[0], w.permute(1, 3, 4, 2, 0).contiguous(), (d // 2, h // 2, wd // 2)
        fn = _lib.lib().gens_conv3d_k3s2
    y = torch.full((1, c_out) + shape, float("nan"), device=DEV)
    st = torch.zeros(2 * c_out, device=DEV, dtype=torch.float64)
    _lib.check(fn(_lib.ptr(x), _lib.ptr(halo) if halo is not None else ctypes.c_void_p(0), _lib.ptr(pk), c_in, c_out, d, h,
                  wd, _lib.ptr(y), _lib.ptr(st), _lib.stream_ptr(DEV)), "strided K13")
    torch.cuda.synchronize()
    return y, st


@pytest.mark.parametrize("transposed,c_in,c_out,shape", [
    (False, 8, 8, (16, 24, 64)), (False, 16, 16, (8, 16, 32)), (False, 24, 8, (4, 6, 10)), (False, 8, 8, (2, 2, 2)),
    (True, 8, 8, (8, 12, 32)), (True, 16, 8, (4, 8, 16)), (True, 8, 8, (3, 5, 7)), (True, 32, 8, (1, 1, 1))])
def test_k13_strided_and_transposed_match_float64(cuda_lib, transposed, c_in, c_out, shape):
    g = torch.Generator().manual_seed(c_in * 7 + c_out + int(transposed))
    x = torch.randn(1, c_in, *shape, generator=g).to(DEV)
    if transposed:
        w = (torch.randn(c_in, c_out, 3, 3, 3, generator=g) / (8 * c_in) ** 0.5).to(DEV)
        want = F.conv_transpose3d(x.double(), w.double(), None, stride=2, padding=1, output_padding=1)
    else:
        w = (torch.randn(c_out, c_in, 3, 3, 3, generator=g) / (27 * c_in) ** 0.5).to(DEV)
        want = F.conv3d(x.double(), w.double(), None, stride=2, padding=1)
    got, st = _strided(x, w, transposed)
    assert got.shape == want.shape and not torch.isnan(got).any()
    assert torch.all((got.double() - want).abs() <= 1e-6 * want.abs().max() + 1e-5 * want.abs())
    sums = torch.cat([want.sum(dim=(0, 2, 3, 4)), (want * want).sum(dim=(0, 2, 3, 4))])
    assert torch.all((st - sums).abs() <= 1e-5 * sums.abs() + 1e-4 * want[0, 0].numel() ** 0.5)


def test_k13_strided_and_transposed_on_slabs(cuda_lib):
    g = torch.Generator().manual_seed(6)
    full = torch.randn(1, 8, 16, 12, 32, generator=g).to(DEV)
    w = (torch.randn(8, 8, 3, 3, 3, generator=g) / 15).to(DEV)
    down = F.conv3d(full.double(), w.double(), None, stride=2, padding=1)
    up = F.conv_transpose3d(full.double(), w.double(), None, stride=2, padding=1, output_padding=1)
    for a0, a1 in ((0, 8), (8, 16), (4, 10)):
        slab = full[:, :, a0:a1].contiguous()
        lo = full[0, :, a0 - 1].contiguous() if a0 > 0 else None
        hi = full[0, :, a1].contiguous() if a1 < 16 else None
        got, _ = _strided(slab, w, False, lo)
        ref = down[:, :, a0 // 2: a1 // 2]
        assert torch.all((got.double() - ref).abs() <= 1e-6 * ref.abs().max() + 1e-5 * ref.abs()), ("down", a0, a1)
        got, _ = _strided(slab, w, True, hi)
        ref = up[:, :, 2 * a0: 2 * a1]
        assert torch.all((got.double() - ref).abs() <= 1e-6 * ref.abs().max() + 1e-5 * ref.abs()), ("up", a0, a1)


def test_slab_ops_accept_strided_halo_views(cuda_lib):
    """A middle rank of the peer-memory path hands BOTH halo planes as strided views of the neighbours' tensors
    (x[:, :, -1:] / x[:, :, :1]); their contiguous copies must both be alive at the launch (regression: the first copy
    was released before the second was made, and the second took over its block)."""
    import torch.nn as nn
    from gens_b200.reg_network import _Ops, _Unit
    g = torch.Generator().manual_seed(21)
    full = torch.randn(1, 8, 24, 16, 64, generator=g).to(DEV)
    ops = _Ops()
    for unit, ref_fn in ((_Unit(8, 8), lambda w: F.conv3d(full.double(), w.double(), None, padding=1)),):
        unit = unit.to(DEV)
        want = ref_fn(unit.conv.weight)
        a0, a1 = 8, 16
        lo, hi = full[:, :, a0 - 1:a0], full[:, :, a1:a1 + 1]          # non-contiguous (channel stride = 24 planes)
        assert not lo.is_contiguous() and not hi.is_contiguous()
        got, _ = ops._conv_k13(full[:, :, a0:a1].contiguous(), unit.conv, lo, hi, True)
        torch.cuda.synchronize()
        ref = want[:, :, a0:a1]
        assert torch.all((got.double() - ref).abs() <= 1e-6 * ref.abs().max() + 1e-5 * ref.abs())
    out = nn.Conv3d(8, 4, 3, 1, 1).to(DEV)
    want = F.conv3d(full.double(), out.weight.double(), out.bias.double(), padding=1)[:, :, 8:16]
    got, _ = ops._conv_k13(full[:, :, 8:16].contiguous(), out, full[:, :, 7:8], full[:, :, 16:17], False)
    torch.cuda.synchronize()
    assert torch.all((got.double() - want).abs() <= 1e-6 * want.abs().max() + 1e-5 * want.abs())


def test_instnorm_relu_kernel(cuda_lib):
    from gens_b200 import _lib
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(1, 8, 8, 16, 32, generator=g) * 3 + 1).to(DEV)
    skip = torch.randn(1, 8, 8, 16, 32, generator=g).to(DEV)
    want = F.relu(F.instance_norm(x.double(), eps=1e-5)) + skip.double()
    st = torch.cat([x.double().sum(dim=(0, 2, 3, 4)), (x.double() ** 2).sum(dim=(0, 2, 3, 4))])
    y = x.clone()
    n = x[0, 0].numel()
    _lib.check(_lib.lib().gens_instnorm_relu(_lib.ptr(y), _lib.ptr(st), 8, n, float(n), 1e-5, _lib.ptr(skip),
                                             _lib.stream_ptr(DEV)), "gens_instnorm_relu")
    torch.cuda.synchronize()
    assert torch.all((y.double() - want).abs() <= 1e-6 + 1e-5 * want.abs())


def _err(outs, ref64):
    return max(float(((a.double() - b).abs() / (1e-5 * b.abs().max() + 1e-4 * b.abs())).max()) for a, b in zip(outs, ref64))


def test_regulariser_cuda_path_is_as_accurate_as_the_library_path(cuda_lib):
    """RegNetwork.forward under no_grad on CUDA (K13 + in-place norm) against a float64 evaluation of the same network,
    next to the reference's own op sequence in fp32 (cuDNN + ATen instance_norm, what forward runs with autograd on);
    and, when the staged reference is present, against the reference module itself."""
    from gens_b200.reg_network import RegNetwork
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(2)
        net = RegNetwork().to(DEV).eval()
        dims = [128, 64, 32, 16, 8]
        vols = [torch.randn(1, 8, d, d, d, device=DEV) for d in dims]
        with torch.no_grad():
            ours = net(vols)
            ref64 = copy.deepcopy(net).double()([v.double() for v in vols])
        lib_path = net(vols)  # autograd enabled: the library op sequence
        assert lib_path[0].requires_grad and not ours[0].requires_grad
        e_ours, e_lib = _err(ours, ref64), _err([o.detach() for o in lib_path], ref64)
        print(f"error vs float64 in units of (1e-5 max + 1e-4 |ref|): K13 path {e_ours:.3f}, library path {e_lib:.3f}")
        assert e_ours <= max(1.0, 1.5 * e_lib)
        import os, sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        path = os.path.join(root, "baseline", "_ref", "GenS", "models", "modules", "reg_network.py")
        if os.path.exists(path):
            import importlib.util
            from gens_b200.config import gens_model_conf
            spec = importlib.util.spec_from_file_location("_ref_reg_network_gpu", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            ref = mod.RegNetwork(gens_model_conf()["reg_network"]).to(DEV).eval()
            ref.load_state_dict(net.state_dict())
            with torch.no_grad():
                e_ref = _err(ref(vols), ref64)
            print(f"the unmodified reference module on this GPU: {e_ref:.3f}")
            assert e_ours <= max(1.0, 1.5 * e_ref)
    finally:
        torch.backends.cudnn.allow_tf32 = old
