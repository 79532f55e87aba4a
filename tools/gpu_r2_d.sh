#!/bin/bash
# round 2, visit D: GPU suite (fixed tests) + K1 walk-x / occupancy variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rP --durations=5 > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
grep -E "passed|failed|^FAILED" gpurun_out/r2d_pytest.log | tail -8
python tools/sweep_k1.py 3 0c,30c,31c,32c,33c,34c,35c,0c > gpurun_out/r2d_sweep.txt 2>&1
python tools/sweep_k1.py 5 0c,30c,32c,35c >> gpurun_out/r2d_sweep.txt 2>&1
cat gpurun_out/r2d_sweep.txt
