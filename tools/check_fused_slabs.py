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
ok = True
ex = parallel.SlabExchange.get(dims, dev, None)


def cases():
    """the bench scene, another seed with 5 views, cameras inside the volume, and one view looking away"""
    yield sc.features, sc.intrs, sc.c2ws
    s5 = make_scene(480, 640, 5, seed=3, with_images=False).to(dev)
    yield s5.features, s5.intrs, s5.c2ws
    inside = sc.c2ws.clone(); inside[:, :3, 3] *= 0.3
    yield sc.features, sc.intrs, inside
    turned = sc.c2ws.clone()
    turned[1, :3, :3] = turned[1, :3, :3] @ torch.tensor([[-1.0, 0, 0], [0, 1, 0], [0, 0, -1.0]], device=dev)
    yield sc.features, sc.intrs, turned


for it, (feats, intrs, c2ws) in enumerate(list(cases()) * 2):  # every case lands in both exchange buffers
    ref_v, ref_m = vol.agg_mean_var(feats, intrs, c2ws)
    # poison BOTH exchange buffers on every rank first: a tile nobody writes (neither the owner over NVLink nor the
    # local zero fill) must show up as NaN instead of as the previous build's identical values
    for b in ex.bufs:
        b.fill_(float("nan"))
    torch.cuda.synchronize(); dist.barrier()
    v, m = parallel.fused_sharded_agg_mean_var(vol, feats, intrs, c2ws, rank, world)
    torch.cuda.synchronize()
    for i in range(len(dims)):
        same = torch.equal(v[i], ref_v[i]) and torch.equal(m[i], ref_m[i])
        if not same:
            bad = (v[i] != ref_v[i]) | torch.isnan(v[i])
            print(f"rank {rank} case {it} scale {i}: {int(bad.sum())} volume elements differ "
                  f"({int(torch.isnan(v[i]).sum())} never written)", flush=True)
        ok &= same
    dist.barrier()
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
