#!/bin/bash
# K1 row-group kernel: parity of every variant, then the timing sweep ($1 = variant list, optional).
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_volume_gpu.py -x -q > gpurun_out/pytest_k1.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_k1.log
tail -15 gpurun_out/pytest_k1.log
timeout 300 python tools/sweep_k1.py 3 $1 > gpurun_out/sweep_k1.txt 2>&1
cat gpurun_out/sweep_k1.txt
