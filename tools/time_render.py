"""Times ImplicitSurface.render on config-2-shaped inputs (GPU box only)."""
import sys, time
sys.path.insert(0, '.')
import torch
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
ro, rd = ro.to(dev), rd.to(dev)
nrays = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for chunk in [int(c) for c in (sys.argv[2].split(',') if len(sys.argv) > 2 else ['512', '4096'])]:
    sel = torch.arange(0, ro.shape[0], ro.shape[0] // nrays, device=dev)[:nrays]
    o, d = ro[sel].contiguous(), rd[sel].contiguous()
    def run():
        outs = []
        with torch.no_grad():
            for a, b in zip(o.split(chunk), d.split(chunk)):
                r = surf.render(a, b, sc.near, sc.far, vols, masks, sc.imgs, sc.features, sc.features, sc.intrs, sc.c2ws, 1.0, None)
                outs.append(r['color_fine'])
        return torch.cat(outs)
    run(); torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter(); c = run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'chunk {chunk}: {nrays} rays in {dt*1e3:.1f} ms -> {nrays*128/dt:.3e} ray-samples/s, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB, valid frac {float((c.abs().sum(-1)>0).float().mean()):.2f}', flush=True)
