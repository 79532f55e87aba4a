"""torchrun --nproc-per-node N tools/check_sharded.py : slab-sharded build + in-place all-gather must be
bit-identical to the single-GPU build on every rank; ray-sharded render must equal the unsharded one."""
import os, sys
sys.path.insert(0, '.')
import torch, torch.distributed as dist
from gens_b200 import parallel
from gens_b200.synthetic import make_scene
from gens_b200.volume import Volume

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl')
dev = torch.device('cuda', local)
sc = make_scene(480, 640, 3, seed=0).to(dev)
vol = Volume(volume_dims=[256, 128, 64, 32, 16])
ref_v, ref_m = vol.agg_mean_var(sc.features, sc.intrs, sc.c2ws)
got_v, got_m = parallel.sharded_agg_mean_var(vol, sc.features, sc.intrs, sc.c2ws, rank, world)
ok = all(torch.equal(a, b) for a, b in zip(ref_v + ref_m, got_v + got_m))
vol2 = Volume(volume_dims=[30, 12])   # not divisible by the world size: staged gather path
r2 = vol2.agg_mean_var(sc.features, sc.intrs, sc.c2ws)
g2 = parallel.sharded_agg_mean_var(vol2, sc.features, sc.intrs, sc.c2ws, rank, world)
ok &= all(torch.equal(a, b) for a, b in zip(r2[0] + r2[1], g2[0] + g2[1]))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print('sharded build bit-identical on all ranks:', bool(flag.item()))
dist.barrier(); dist.destroy_process_group()
