"""Where the whole-volume RegNetwork spends its time on one GPU (256^3 pyramid, fp32, TF32 off)."""
import sys
sys.path.insert(0, '.')
import torch
import torch.nn.functional as F
from gens_b200.reg_network import RegNetwork
dev = torch.device('cuda:0')
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = RegNetwork().to(dev).eval()
dims = [256, 128, 64, 32, 16]
vols = [torch.randn(1, 8, d, d, d, device=dev) for d in dims]

def t(fn, n=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

with torch.no_grad():
    from gens_b200.reg_network import _LocalOps
    print(f"whole network, library op sequence (cuDNN + ATen instance_norm): {t(lambda: net._run(vols, _LocalOps)):.2f} ms")
    print(f"whole network, K13 path: {t(lambda: net(vols)):.2f} ms")
    import ctypes
    from gens_b200 import _lib
    for cin, cout in ((8, 8), (8, 4)):
        xx = torch.randn(1, cin, 256, 256, 256, device=dev)
        ww = torch.randn(cout, cin, 3, 3, 3, device=dev).permute(1, 3, 4, 2, 0).contiguous()
        yy = torch.empty(1, cout, 256, 256, 256, device=dev)
        st = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
        null = ctypes.c_void_p(0)
        f = lambda: _lib.lib().gens_conv3d_k3(_lib.ptr(xx), null, null, _lib.ptr(ww), null, cin, cout, 256, 256, 256, _lib.ptr(yy), _lib.ptr(st), _lib.stream_ptr(dev))
        if cout == 8:
            _lib.lib().gens_debug_conv_td8(1)
            print(f"K13 conv 8->8 with 8-voxel columns: {t(f, 5):.3f} ms")
            _lib.lib().gens_debug_conv_td8(0)
        ms = t(f, 5)
        print(f"K13 conv {cin}->{cout} @256^3: {ms:.3f} ms = {2 * 27 * cin * cout * 256 ** 3 / ms / 1e9:.1f} TFLOP/s fp32, {(cin + cout) * 256 ** 3 * 4 / ms / 1e6:.0f} GB/s algorithmic")
        del xx, yy
    x = vols[0]
    w = net.conv0.conv.weight
    y = F.conv3d(x, w, None, 1, 1)
    print(f"conv0 8->8 @256^3: {t(lambda: F.conv3d(x, w, None, 1, 1)):.2f} ms")
    print(f"F.instance_norm @256^3 x 8 ch: {t(lambda: F.instance_norm(y, eps=1e-5)):.2f} ms")
    def mine():
        var, mean = torch.var_mean(y, dim=(2, 3, 4), unbiased=False)
        return ((y - mean.view(1, -1, 1, 1, 1)) * torch.rsqrt(var + 1e-5).view(1, -1, 1, 1, 1)).relu_()
    print(f"var_mean + normalise + relu: {t(mine):.2f} ms")
    w2 = net.encoder_layers[0][0].conv.weight
    print(f"enc0 stride-2 8->8: {t(lambda: F.conv3d(y, w2, None, 2, 1)):.2f} ms")
    z = torch.randn(1, 8, 128, 128, 128, device=dev)
    wd = net.decoder_layers[0].conv.weight
    print(f"dec0 transposed 8->8 128->256: {t(lambda: F.conv_transpose3d(z, wd, None, 2, 1, 1)):.2f} ms")
    wo = net.out_layers[0]
    print(f"out0 8->4 @256^3: {t(lambda: F.conv3d(y, wo.weight, wo.bias, 1, 1)):.2f} ms")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net(vols); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
