#!/bin/bash
# round 2, visit C (2 GPUs): fused slab exchange with culling (bit-identity incl. poisoned buffers), 2-GPU tests, bench N=2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  tools/check_fused_slabs.py > gpurun_out/r2c_fused.log 2>&1
echo "check_fused rc=$?"; grep -E "bit-identical|world|differ" gpurun_out/r2c_fused.log | head -20; tail -3 gpurun_out/r2c_fused.log
timeout 600 python -m pytest tests -m gpu -q -k "two_gpus" -rP > gpurun_out/r2c_pytest2.log 2>&1
echo "pytest2 rc=$?"; tail -5 gpurun_out/r2c_pytest2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
echo "bench n2 rc=$?"; cut -c1-3000 gpurun_out/r2c_bench_n2.json; tail -5 gpurun_out/r2c_bench_n2.err
