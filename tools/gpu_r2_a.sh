#!/bin/bash
# round 2, visit A: full GPU suite (incl. drop-in proof vs the staged reference), K1 variant sweep, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
python -m pytest tests -m gpu -q -rs --durations=15 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -40 gpurun_out/r2a_pytest.log
python tools/sweep_k1.py 3 10,11,12c,13,0,0c > gpurun_out/r2a_sweep3.txt 2>&1
python tools/sweep_k1.py 5 11,0,0c >> gpurun_out/r2a_sweep3.txt 2>&1
cat gpurun_out/r2a_sweep3.txt
python bench.py --steps 20 --warmup 5 --no-render > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
