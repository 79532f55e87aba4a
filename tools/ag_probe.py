import os, sys, time
sys.path.insert(0, '.')
import torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dist.init_process_group('nccl'); dev = torch.device('cuda', local)
dims = [256, 128, 64, 32, 16]
fulls = [torch.zeros(1, 9, d, d, d, device=dev) for d in dims]
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def inplace(coalesce):
    def f():
        if coalesce:
            with dist._coalescing_manager(device=dev, async_ops=False):
                for t, d in zip(fulls, dims):
                    p = d // world
                    for c in range(9): dist.all_gather_into_tensor(t[0, c], t[0, c, rank * p:(rank + 1) * p])
        else:
            for t, d in zip(fulls, dims):
                p = d // world
                for c in range(9): dist.all_gather_into_tensor(t[0, c], t[0, c, rank * p:(rank + 1) * p])
    return f
tot = sum(9 * d ** 3 for d in dims)
send = torch.zeros(tot // world, device=dev); recv = torch.zeros(tot, device=dev)
def flat():
    dist.all_gather_into_tensor(recv, send)
big = torch.zeros(9 * 256 ** 3, device=dev)
def one_big():
    dist.all_gather_into_tensor(big, big[rank * (big.numel() // world):(rank + 1) * (big.numel() // world)])
r = {'inplace45_coalesced': timeit(inplace(True)), 'inplace45_plain': timeit(inplace(False)), 'flat_single': timeit(flat), 'one_big_inplace': timeit(one_big)}
if rank == 0: print(world, {k: round(v, 3) for k, v in r.items()}, 'ms; payload per rank MB', tot * 4 / world / 1e6)
dist.barrier(); dist.destroy_process_group()
