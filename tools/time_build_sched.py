"""Device time of the 5-scale build for the build-level scheduling variants (GPU box only)."""
import sys
sys.path.insert(0, '.')
import torch
from gens_b200 import _lib
from gens_b200.synthetic import make_scene
from gens_b200.volume import Volume

dev = torch.device('cuda:0')
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 3
scheds = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 1, 2, 3, 7]
sc = make_scene(480, 640, nv, seed=0, with_images=False).to(dev)
vol = Volume(volume_dims=[256, 128, 64, 32, 16])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib()
ref = None
for s in scheds:
    L.gens_debug_set_variant(100 + s)
    for _ in range(5):
        out = vol.agg_mean_var(sc.features, sc.intrs, sc.c2ws)
    ts = []
    for _ in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = vol.agg_mean_var(sc.features, sc.intrs, sc.c2ws); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    same = True
    if ref is None:
        ref = [[t.clone() for t in o] for o in out]
    else:
        same = all(torch.equal(x, y) for o, r in zip(out, ref) for x, y in zip(o, r))
    print(f"sched {s}: median {ts[15]:.1f} us, min {ts[0]:.1f} us{'' if same else '  MISMATCH'}", flush=True)
L.gens_debug_set_variant(103)
