"""world_size-2 gloo test of the sharding logic: slab-built volumes gathered over the process group are
bit-identical to the full build, and ray shards re-assemble in order.  Slabs are produced by the C oracle
(the CUDA kernel needs a GPU); the host-side partitioning / collective code is the product's."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gens_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, golden_dir, out_q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import c_oracle
        g = np.load(f"{golden_dir}/volume_agg.npz")
        w2c = torch.inverse(torch.from_numpy(g["c2ws"])).numpy()
        ok = True
        for i, d in enumerate([32, 16, 6]):
            d = int(d)
            feat = g[f"feat{min(i, 4)}"]
            k = torch.from_numpy(g["intrs"]).clone()
            k[:, :2] *= 0.5 ** min(i, 4)
            grid = torch.linspace(-1, 1, d).numpy()
            full_v, full_m = c_oracle.volume_agg(feat, w2c, k.numpy(), grid)
            a0, a1 = parallel.slab_bounds(d, rank, world)
            v, m = c_oracle.volume_agg(feat, w2c, k.numpy(), grid, a0=a0, a1=a1)
            slab_v = torch.from_numpy(v[None, :, a0:a1].copy())
            slab_m = torch.from_numpy(m[None, None, a0:a1].copy())
            got_v = parallel.gather_slabs(slab_v, d, world)
            got_m = parallel.gather_slabs(slab_m, d, world)
            ok &= np.array_equal(got_v[0].numpy(), full_v) and np.array_equal(got_m[0, 0].numpy(), full_m)
        # in-place per-channel gather (the NCCL path of sharded_agg_mean_var): slab already sits in the full tensor
        d = 8
        ref = torch.arange(3 * d ** 3, dtype=torch.float32).reshape(1, 3, d, d, d)
        full = torch.full_like(ref, -1.0)
        a0, a1 = parallel.slab_bounds(d, rank, world)
        full[:, :, a0:a1] = ref[:, :, a0:a1]
        parallel.gather_slabs_inplace(full, d, rank, world)
        ok &= torch.equal(full, ref)
        n = 1001
        lo, hi = parallel.shard_range(n, rank, world)
        rays = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3)
        got = parallel.gather_rays(rays[lo:hi] * 2, n, rank, world)
        ok &= torch.equal(got, rays * 2)
        # mesh-extraction lattice sharded by x-slabs (config 5): a stub evaluator stands in for the CUDA MLP
        res = 21  # not divisible by the world size: shards differ by one plane
        lattice = torch.arange(res ** 3, dtype=torch.float32).reshape(res, res, res).sin()
        calls = []

        def sdf_grid_fn(x_range):
            calls.append(tuple(x_range))
            return lattice[x_range[0]:x_range[1]].clone()
        on_dst = parallel.sharded_sdf_grid(sdf_grid_fn, res, rank, world, dst=0)
        ok &= calls == [parallel.shard_range(res, rank, world)]
        ok &= (on_dst is None) if rank != 0 else torch.equal(on_dst, lattice)
        everywhere = parallel.sharded_sdf_grid(sdf_grid_fn, res, rank, world, dst=None)
        ok &= torch.equal(everywhere, lattice)
        # sharded mesh extraction: every rank meshes its x-slab (+ one plane of overlap), only the meshes travel.  The
        # CPU oracle's sequential mesher stands in for the CUDA kernels; the merged mesh must be the full lattice's.
        from gens_b200 import meshing
        from oracle import mc_oracle
        gx = np.linspace(-1, 1, 17)
        xx, yy, zz = np.meshgrid(gx, gx, gx, indexing="ij")
        field = (np.sqrt(xx * xx + yy * yy + zz * zz) - 0.6).astype(np.float32)

        def cpu_mesher(u, iso, index_offset=(0.0, 0.0, 0.0)):
            v, t = mc_oracle.marching_cubes_numpy(u.numpy(), iso)
            return torch.from_numpy(v + np.asarray(index_offset)[None, :]), torch.from_numpy(t)
        x0, x1 = parallel.shard_range(17, rank, world)
        slab = torch.from_numpy(field[x0:min(x1 + 1, 17)])
        mv, mt = meshing.sharded_marching_cubes(slab, 0.0, x0, rank, world, dst=None, mesher=cpu_mesher)
        fv, ft = mc_oracle.marching_cubes_numpy(field, 0.0)
        # this CPU stand-in adds the slab offset AFTER the interpolation ((i_local + t) + x0, one ulp away from
        # (i_global + t)); the CUDA mesher adds it to the integer index first and is exact (test_marching_cubes_gpu.py)
        merged, whole = mc_oracle.canonical_triangles(mv.numpy(), mt.numpy()), mc_oracle.canonical_triangles(fv, ft)
        ok &= merged.shape == whole.shape and bool(np.allclose(merged, whole, rtol=0.0, atol=1e-12))
        only0 = meshing.sharded_marching_cubes(slab, 0.0, x0, rank, world, dst=0, mesher=cpu_mesher)
        ok &= (only0[0] is None) if rank != 0 else (only0[1].shape[0] == ft.shape[0])
        out_q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_slab_and_ray_sharding_world2(golden_dir):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, golden_dir, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {0: True, 1: True}


def test_bounds_cover_everything():
    for d in (16, 30, 256):
        for world in (1, 2, 3, 8):
            edges = [parallel.slab_bounds(d, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == d
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
