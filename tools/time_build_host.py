"""Host-side cost of one Volume.agg_mean_var call (GPU box only): where does the interpreter time go?"""
import sys, time
sys.path.insert(0, '.')
import torch
from gens_b200.synthetic import make_scene
from gens_b200.volume import Volume
import gens_b200.volume as V

dev = torch.device('cuda:0')
sc = make_scene(480, 640, 3, seed=0, with_images=False).to(dev)
vol = Volume(volume_dims=[256, 128, 64, 32, 16])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    out = vol.agg_mean_var(sc.features, sc.intrs, sc.c2ws)
torch.cuda.synchronize()

def cpu_time(fn, n=200, sync_each=False):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        o = fn()
        ts.append(time.perf_counter() - t0)
        if sync_each:
            torch.cuda.synchronize()
        del o
    torch.cuda.synchronize()
    ts.sort()
    return ts[len(ts) // 2] * 1e6, ts[-1] * 1e6

print("agg_mean_var, GPU idle at call (sync each)  : median %.1f us, max %.1f us" % cpu_time(lambda: vol.agg_mean_var(sc.features, sc.intrs, sc.c2ws), sync_each=True))
print("agg_mean_var, back to back (queue may fill) : median %.1f us, max %.1f us" % cpu_time(lambda: vol.agg_mean_var(sc.features, sc.intrs, sc.c2ws)))
print("5 x torch.empty of the output sizes         : median %.1f us, max %.1f us" % cpu_time(lambda: [torch.empty((1, 9, d, d, d), device=dev) for d in (256, 128, 64, 32, 16)], sync_each=True))
print("flush.zero_()                               : median %.1f us, max %.1f us" % cpu_time(lambda: flush.zero_(), sync_each=True))
print("event pair create+record                    : median %.1f us, max %.1f us" % cpu_time(lambda: [torch.cuda.Event(enable_timing=True).record(), torch.cuda.Event(enable_timing=True).record()], sync_each=True))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    o = vol.agg_mean_var(sc.features, sc.intrs, sc.c2ws)
    torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
# device-side duration of one call with the GPU idle at the start vs kept busy
for label, pre in (("GPU idle at a.record", lambda: torch.cuda.synchronize()), ("behind a 256 MiB memset", lambda: flush.zero_())):
    ts = []
    for _ in range(30):
        pre()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); o = vol.agg_mean_var(sc.features, sc.intrs, sc.c2ws); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"device time of one build, {label}: median {ts[15]:.1f} us, min {ts[0]:.1f} us")
