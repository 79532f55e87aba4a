#!/bin/bash
# round 2, visit I (8 GPUs): fused slab exchange at 8 ranks (bit-identity, poisoned buffers), bench at N = 8 and N = 4
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  tools/check_fused_slabs.py > gpurun_out/r2i_fused8.log 2>&1
echo "check_fused8 rc=$?"; grep -E "bit-identical|world|differ" gpurun_out/r2i_fused8.log | sort | uniq -c | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n8.json 2> gpurun_out/r2i_bench_n8.err
echo "bench n8 rc=$?"; cut -c1-1800 gpurun_out/r2i_bench_n8.json; tail -3 gpurun_out/r2i_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 \
  bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n4.json 2> gpurun_out/r2i_bench_n4.err
echo "bench n4 rc=$?"; cut -c1-1200 gpurun_out/r2i_bench_n4.json
