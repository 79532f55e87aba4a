#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list of the bench command.
# Everything lands under gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 1500 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --render-steps 1 \
  > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit $?"
wc -l gpurun_out/launches_bench.csv
