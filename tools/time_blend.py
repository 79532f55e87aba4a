"""K10 timing under both weight sources (shared memory / constant bank) + agreement with the module forward."""
import sys
sys.path.insert(0, '.')
import torch
from gens_b200 import _lib
from gens_b200.networks import BlendingNetwork
dev = torch.device('cuda:0')
torch.manual_seed(0)
net = BlendingNetwork(d_feature=20).to(dev)
for ns in (2, 4):
    n = 1 << 22
    rf = torch.rand(n, ns, 23, device=dev)
    rd = torch.randn(n, ns, 4, device=dev) * 0.3
    m = torch.rand(n, ns, device=dev) > 0.2
    with torch.no_grad():
        ref = net(rf[:65536], rd[:65536], m[:65536])
        for knob in (0, 1):
            _lib.lib().gens_debug_blend_const(knob)
            for _ in range(2):
                out = net.blend_nograd(rf, rd, m)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                out = net.blend_nograd(rf, rd, m)
            b.record(); torch.cuda.synchronize()
            print(f"ns={ns} const={knob}: {a.elapsed_time(b) / 3:.3f} ms per {n} points, max |diff| vs module "
                  f"{float((out[:65536] - ref).abs().max()):.2e}", flush=True)
_lib.lib().gens_debug_blend_const(1)
