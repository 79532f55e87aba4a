#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rP --durations=5 > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
grep -E "passed|failed|^FAILED" gpurun_out/r2e_pytest.log | tail -8
python tools/sweep_k1.py 3 0c,36c,37c,38c,39c,0c > gpurun_out/r2e_sweep.txt 2>&1
cat gpurun_out/r2e_sweep.txt
