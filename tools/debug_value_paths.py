"""GPU debug: SDF value paths (autograd forward, fp32 fused chain, tensor-core kernel) against each other."""
import sys
sys.path.insert(0, '.')
import torch
from gens_b200 import sdf_analytic
from gens_b200.config import gens_model_conf
from gens_b200.implicit_surface import ImplicitSurface
from gens_b200.synthetic import make_reg_volumes
dev = torch.device('cuda:0')
torch.manual_seed(0)
surf = ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"]).to(dev)
for dims in ([32, 16, 8, 4, 2], [64, 32, 16, 8, 4]):
    vols = [v.to(dev) for v in make_reg_volumes(dims, seed=3)]
    for n in (1920, 16384, 65536, 262144):
        pts = torch.rand(n, 3, device=dev) * 2.2 - 1.1
        with torch.no_grad():
            ref = surf.sdf_network(pts, vols)[:, :1]
            out = {}
            for tc in (False, True):
                sdf_analytic.USE_TC = tc
                out[tc] = surf.sdf_network.sdf_nograd(pts, vols)
            sdf_analytic.USE_TC = True
        print(dims[0], n, 'fp32 chain vs forward', float((out[False] - ref).abs().max()), 'tc vs forward',
              float((out[True] - ref).abs().max()), flush=True)
