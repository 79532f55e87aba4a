"""Config-3 training step (bench.py's train leg): wall time per step vs summed kernel time and launch count."""
import sys
sys.path.insert(0, '.')
import torch
import bench
from gens_b200.losses import compute_LNCC
from gens_b200.synthetic import make_scene
from gens_b200.volume import Volume
dev = torch.device('cuda:0')
host = make_scene(480, 640, 5, seed=0, with_images=True)
sc = host.to(dev)
surf = bench.build_surface(dev); surf.train()
vols = [v.requires_grad_(True) for v in bench.smooth_volumes(bench.DIMS, dev)]
feats = [f.clone().requires_grad_(True) for f in sc.features]
with torch.no_grad():
    _, masks = Volume(volume_dims=bench.DIMS).agg_mean_var(sc.features, sc.intrs, sc.c2ws)
step = bench.train_step_fn(surf, sc, vols, masks, feats, 512, compute_LNCC, dev)
for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"wall {1e3 * (time.perf_counter() - t0) / 5:.1f} ms per step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev) if hasattr(ev[0], 'device_time') else sum(e.cuda_time for e in ev)
print(f"{len(ev)} device events, {tot / 1e3:.1f} ms of device time in one step")
print(prof.key_averages().table(sort_by="cuda_time_total" if not hasattr(ev[0], 'device_time') else "device_time_total", row_limit=25, max_name_column_width=60))
