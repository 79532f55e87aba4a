"""Runs the tcgen05 SDF value kernel a few times on random encodings (for ncu; GPU box only)."""
import sys
sys.path.insert(0, '.')
import torch
from gens_b200 import mlp_tc
from gens_b200.config import gens_model_conf
from gens_b200.implicit_surface import ImplicitSurface
from gens_b200.sdf_analytic import FoldedSDF
dev = torch.device('cuda:0')
torch.manual_seed(0)
surf = ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"]).to(dev)
packed = mlp_tc.PackedSDF(FoldedSDF(surf.sdf_network))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
pos = torch.randn(n, 27, device=dev).clamp(-1, 1)
fe = torch.randn(n, 100, device=dev).clamp(-1, 1)
for _ in range(3):
    out = mlp_tc.sdf_values(packed, pos, fe)
torch.cuda.synchronize()
print(float(out.mean()))
