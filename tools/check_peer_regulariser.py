"""torchrun --nproc-per-node N tools/check_peer_regulariser.py : the slab regulariser through NVLink peer memory (eager and
as a CUDA graph) against the NCCL slab path, per rank and per scale.  GPU box only."""
import os, sys
sys.path.insert(0, '.')
import torch
import torch.distributed as dist
from gens_b200.synthetic import make_scene
from gens_b200.volume import Volume
from gens_b200.config import gens_model_conf
from gens_b200.reg_network import RegNetwork, PeerSlabRegulariser
from gens_b200 import parallel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world)
torch.backends.cudnn.allow_tf32 = False
dims = [256, 128, 64, 32, 16]
sc = make_scene(480, 640, 3, seed=0, with_images=False).to(dev)
vol = Volume(volume_dims=dims)
torch.manual_seed(0)
net = RegNetwork(gens_model_conf()["reg_network"]).to(dev).eval()


def diffs(a, b):
    return [f"{float((x - y).abs().max() / y.abs().max()):.1e}" for x, y in zip(a, b)]


with torch.no_grad():
    ref_v, ref_m = parallel.sharded_build_and_regularise(vol, net, sc.features, sc.intrs, sc.c2ws, rank, world)
    ref_v = [v.clone() for v in ref_v]
    for mode in ("eager", "graph"):
        peer = PeerSlabRegulariser(net, dims, rank, world, dev, use_graph=(mode == "graph"))
        for it in range(3):
            v, m = parallel.sharded_build_and_regularise(vol, net, sc.features, sc.intrs, sc.c2ws, rank, world, graphed=peer)
            torch.cuda.synchronize()
            print(f"rank {rank} {mode} (graph={peer.graph is not None}) run {it}: vs NCCL path {diffs(v, ref_v)} masks "
                  f"{all(torch.equal(a, b) for a, b in zip(m, ref_m))}", flush=True)
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); a.record()
        for _ in range(5):
            parallel.sharded_build_and_regularise(vol, net, sc.features, sc.intrs, sc.c2ws, rank, world, graphed=peer)
        b.record(); torch.cuda.synchronize()
        if rank == 0:
            print(f"world {world} peer {mode}: {a.elapsed_time(b) / 5:.2f} ms per step", flush=True)
        del peer
dist.barrier(); dist.destroy_process_group()
