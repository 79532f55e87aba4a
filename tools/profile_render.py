"""torch.profiler breakdown of one ImplicitSurface.render call (GPU box only)."""
import sys
sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
from gens_b200.config import gens_model_conf
from gens_b200.implicit_surface import ImplicitSurface
from gens_b200.synthetic import make_scene
from gens_b200.volume import Volume

dev = torch.device('cuda:0')
dims = [256, 128, 64, 32, 16]
sc = make_scene(480, 640, 3, seed=0).to(dev)
torch.manual_seed(0)
surf = ImplicitSurface(gens_model_conf(perturb=1.0)["implicit_surface"]).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
vols = []
for d in dims:
    base = torch.randn(1, 4, max(d // 8, 2), max(d // 8, 2), max(d // 8, 2), device=dev, generator=g) * 0.5
    vols.append(torch.nn.functional.interpolate(base, size=(d, d, d), mode='trilinear', align_corners=True).contiguous())
_, masks = Volume(volume_dims=dims).agg_mean_var(sc.features, sc.intrs, sc.c2ws)
ro, rd = sc.rays(step=1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sel = torch.arange(0, ro.shape[0], ro.shape[0] // n, device=dev)[:n]
o, d = ro[sel].contiguous(), rd[sel].contiguous()
def run():
    with torch.no_grad():
        return surf.render(o, d, sc.near, sc.far, vols, masks, sc.imgs, sc.features, sc.features, sc.intrs, sc.c2ws, 1.0, None)
run(); run(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=int(sys.argv[2]) if len(sys.argv) > 2 else 45, max_name_column_width=70))
