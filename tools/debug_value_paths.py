"""GPU debug: SDF value paths (autograd forward, fp32 fused chain, tensor-core kernel) against each other."""
import sys
sys.path.insert(0, '.')
import torch
from gens_b200 import sdf_analytic
from gens_b200.config import gens_model_conf
from gens_b200.implicit_surface import ImplicitSurface
from gens_b200.synthetic import make_reg_volumes
from gens_b200 import _lib
import copy
dev = torch.device('cuda:0')
torch.manual_seed(0)
surf = ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"]).to(dev)
for dims in ([32, 16, 8, 4, 2], [64, 32, 16, 8, 4]):
    vols = [v.to(dev) for v in make_reg_volumes(dims, seed=3)]
    for n in (1920, 16384, 65536, 262144):
        pts = torch.rand(n, 3, device=dev) * 2.2 - 1.1
        with torch.no_grad():
            ref = surf.sdf_network(pts, vols)[:, :1]
            # float64 reference: same module in double on the fp32 look-up (the look-up is not what is being compared)
            feats = surf.sdf_network._lookup(pts, vols).double()
            net64 = copy.deepcopy(surf.sdf_network).double()
            net64._lookup = lambda p, v: feats
            ref64 = net64(pts.double(), vols)[:, :1]
            out = {}
            for name, tc, terms in (("fp32 chain", False, 3), ("tc 3 terms", True, 3), ("tc 4 terms", True, 4)):
                sdf_analytic.USE_TC = tc
                _lib.lib().gens_debug_set_tc_terms(terms)
                out[name] = surf.sdf_network.sdf_nograd(pts, vols)
            sdf_analytic.USE_TC = True
            _lib.lib().gens_debug_set_tc_terms(3)
        line = f"D{dims[0]} n={n}: autograd fp32 forward vs f64 {float((ref - ref64).abs().max()):.2e}"
        for name, v in out.items():
            line += f" | {name} vs f64 {float((v - ref64).abs().max()):.2e}"
        print(line, flush=True)
