#!/usr/bin/env python
"""bench.py -- headline benchmark of the GenS hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over BASELINE config 2 (480x640, 3 views, volume dims
[256,128,64,32,16]): the 5-scale volume build.  `value` = voxel*views/s with inputs resident in
HBM; `e2e` = the same through the public API with pinned host buffers, H2D of the feature
pyramid + cameras and D2H of the volumes inside the timed region.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = [256, 128, 64, 32, 16]
HW = (480, 640)
L2_FLUSH_BYTES = 256 << 20


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for n, p in zip(names, parts[3:7]):
                if p.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x: float, world: int, device) -> float:
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def voxel_views(nv):
    return sum(d ** 3 for d in DIMS) * nv


def algorithmic_bytes_scale(d, nv, h, w):
    """8 channels + mask written once, each feature map read once."""
    return d ** 3 * 9 * 4 + nv * 4 * h * w * 4


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path: the same ATen op sequence as
    models/modules/volume.py on host tensors with every host thread (oracle/torch_oracle.py; the
    reference is Python, so there is no oracle/_ref binary -- kind = "port")."""
    if rank != 0:
        return
    from gens_b200.synthetic import make_scene
    from oracle import torch_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc = make_scene(HW[0], HW[1], args.nv, seed=0, with_images=False)
    run = lambda: torch_oracle.agg_mean_var(sc.features, sc.intrs, sc.c2ws, DIMS)
    with torch.no_grad():
        for _ in range(args.warmup):
            run()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run()
        dt = (time.perf_counter() - t0) / args.steps
    val = voxel_views(args.nv) / dt
    line = {
        "impl": "reference", "metric": "voxel*views/s (volume build)", "value": val, "unit": "voxel*views/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"config2: {HW[0]}x{HW[1]}, {args.nv} views, volume dims {DIMS}, full 5-scale build",
                   "device": "host cpu"},
        "cpu_baseline": {"value": val, "unit": "voxel*views/s", "cores": cores, "kind": "port",
                         "sample": "full 5-scale build per step (ATen-op restatement of volume.py on host tensors)"},
        "e2e": {"value": val, "unit": "voxel*views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- gens_b200 has no CPU fallback (use --impl reference)")
    from gens_b200 import _lib, build
    from gens_b200.synthetic import make_scene
    from gens_b200.volume import Volume, stage_cameras, pack_feature_maps
    if rank == 0:
        build.build()
    barrier(world)
    _lib.lib()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    nv = args.nv
    host = make_scene(HW[0], HW[1], nv, seed=0, with_images=False)
    sc = host.to(dev)
    vol_mod = Volume(volume_dims=DIMS)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    # slab sharding across ranks (planes of tensor dim 2); world == 1 -> full build
    def step_device():
        if world == 1:
            return vol_mod.agg_mean_var(sc.features, sc.intrs, sc.c2ws)
        from gens_b200.parallel import sharded_agg_mean_var
        return sharded_agg_mean_var(vol_mod, sc.features, sc.intrs, sc.c2ws, rank, world)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        barrier(world)
        evs = []
        t0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()  # L2 flush, outside the event pair
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            evs.append((a, b))
            del out
        torch.cuda.synchronize(dev)
        barrier(world)
        wall = time.perf_counter() - t0
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        return max_over_ranks(dev_ms, world, dev) / steps, wall

    with ClockSampler(local) as clocks:
        ms_step, wall = timed(step_device, args.steps, args.warmup)

        # ---- roofline of the dominant kernel: K1 on the 256^3 scale, timed alone ----------
        d0 = DIMS[0]
        feat_cl = pack_feature_maps(sc.features[0])
        w2c, k0 = stage_cameras(sc.intrs, sc.c2ws, 0)
        grid0 = torch.linspace(-1, 1, d0, device=dev)
        vol0 = torch.empty((8, d0, d0, d0), device=dev)
        msk0 = torch.empty((d0, d0, d0), device=dev)
        stream = _lib.stream_ptr(dev)

        def k1():
            _lib.check(_lib.lib().gens_volume_agg_fwd(
                _lib.ptr(feat_cl), nv, HW[0], HW[1], _lib.ptr(w2c), _lib.ptr(k0), 1.0, _lib.ptr(grid0), d0, 0, d0, 0,
                d0 ** 3, 1, _lib.DIV_RECIP, _lib.ptr(vol0), _lib.ptr(msk0), stream), "K1")
        k1_ms, _ = timed(k1, max(args.steps, 10), 3)
        del vol0, msk0

        # ---- e2e: pinned host buffers -> public API -> pinned host result ------------------
        pin_feats = [f.pin_memory() for f in host.features]
        pin_intrs, pin_c2ws = host.intrs.pin_memory(), host.c2ws.pin_memory()
        pin_out = [torch.empty((1, 9, d, d, d), dtype=torch.float32).pin_memory() for d in DIMS]
        h2d = sum(f.numel() for f in pin_feats) * 4 + (pin_intrs.numel() + pin_c2ws.numel()) * 4
        d2h = sum(o.numel() for o in pin_out) * 4

        def step_e2e():
            feats = [f.to(dev, non_blocking=True) for f in pin_feats]
            vols, masks = vol_mod.agg_mean_var(feats, pin_intrs.to(dev, non_blocking=True),
                                               pin_c2ws.to(dev, non_blocking=True))
            for o, v, m in zip(pin_out, vols, masks):
                o[:, :8].copy_(v, non_blocking=True)
                o[:, 8:].copy_(m, non_blocking=True)
        e2e_ms = None
        if world == 1:
            e2e_ms, _ = timed(step_e2e, max(3, args.steps // 4), 2)
    clk = clocks.summary()

    fill = None
    if rank == 0:
        _, masks = vol_mod.agg_mean_var(sc.features, sc.intrs, sc.c2ws)
        fill = [round(m.mean().item(), 4) for m in masks]

    # ---- CPU baseline (rank 0, N=1): bounded sample of the same workload --------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import torch_oracle
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        with torch.no_grad():
            torch_oracle.agg_mean_var(host.features, host.intrs, host.c2ws, DIMS)
            best = 1e30
            for _ in range(2):
                t0 = time.perf_counter()
                torch_oracle.agg_mean_var(host.features, host.intrs, host.c2ws, DIMS)
                best = min(best, time.perf_counter() - t0)
        cpu = {"value": voxel_views(nv) / best, "unit": "voxel*views/s", "cores": cores, "kind": "port",
               "sample": "full 5-scale build, best of 2 after 1 warm-up (ATen-op restatement of volume.py, "
                         "all host threads)", "ms": best * 1e3}

    if rank != 0:
        return
    peak, peak_kind = measured_peak_gbs()
    abytes = algorithmic_bytes_scale(d0, nv, HW[0], HW[1])
    achieved = abytes / (k1_ms * 1e-3) / 1e9
    vv = voxel_views(nv)
    line = {
        "metric": "voxel*views/s (volume build)", "value": vv / (ms_step * 1e-3), "unit": "voxel*views/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"config2: {HW[0]}x{HW[1]}, {nv} views, volume dims {DIMS}, full 5-scale build",
                   "sharding": "none" if world == 1 else f"x-slabs over {world} ranks + all-gather",
                   "l2": "256 MiB memset between steps, outside the per-step event pairs",
                   "mask_fill": fill, "wall_s_timed_loop": round(wall, 4)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "volume_agg_packed_kernel @ 256^3", "ms": k1_ms,
                     "algorithmic_bytes": abytes, "peak_source": f"MEASURED_PEAKS.json ({peak_kind}, burst copy)"},
        "cpu_baseline": cpu,
        "e2e": None if e2e_ms is None else {"value": vv / (e2e_ms * 1e-3), "unit": "voxel*views/s",
                                            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                            "ms_per_step": e2e_ms},
        "gpu_launches": 2 * len(DIMS) * args.steps,  # per step: 5 pack + 5 aggregation kernels
        "clocks": clk,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nv", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank, world, local = dist_setup(args.gpus)
    if args.impl == "reference":
        args.warmup = min(args.warmup, 2)  # each step is a full 3-5 s CPU build
        run_reference(args, rank, world)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
