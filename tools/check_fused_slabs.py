"""torchrun --nproc-per-node N tools/check_fused_slabs.py : fused build + NVLink peer stores vs the 1-GPU build
(bit-identical?) and vs the all-gather path (time).  GPU box only."""
import os, sys, time
sys.path.insert(0, '.')
import torch
import torch.distributed as dist
from gens_b200.synthetic import make_scene
from gens_b200.volume import Volume
from gens_b200 import parallel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world)
dims = [256, 128, 64, 32, 16]
sc = make_scene(480, 640, 3, seed=0, with_images=False).to(dev)
vol = Volume(volume_dims=dims)
ref_v, ref_m = vol.agg_mean_var(sc.features, sc.intrs, sc.c2ws)
ok = True
for it in range(3):  # three builds: both buffers, and a reuse
    v, m = parallel.fused_sharded_agg_mean_var(vol, sc.features, sc.intrs, sc.c2ws, rank, world)
    torch.cuda.synchronize()
    for i in range(len(dims)):
        ok &= torch.equal(v[i], ref_v[i]) and torch.equal(m[i], ref_m[i])
print(f"rank {rank}: fused build bit-identical to the 1-GPU build: {ok}", flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, steps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); o = fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b); del o
    t = torch.tensor([tot / steps], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
t_fused = timed(lambda: parallel.fused_sharded_agg_mean_var(vol, sc.features, sc.intrs, sc.c2ws, rank, world))
t_nccl = timed(lambda: parallel.sharded_agg_mean_var(vol, sc.features, sc.intrs, sc.c2ws, rank, world))
t_one = timed(lambda: vol.agg_mean_var(sc.features, sc.intrs, sc.c2ws))
if rank == 0:
    print(f"world {world}: fused build+exchange {t_fused*1e3:.0f} us | slabs + all-gather + scatter {t_nccl*1e3:.0f} us | 1-GPU full build {t_one*1e3:.0f} us", flush=True)
dist.barrier(); dist.destroy_process_group()
