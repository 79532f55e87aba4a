#!/bin/bash
# round 2, visit B: GPU suite again (fixed tests), ncu --set full of the shipped K1 (constant-bank cameras + bulk zero fill)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rP --durations=8 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
grep -E "passed|failed" gpurun_out/r2b_pytest.log | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'volume_agg_rowgroup' -s 3 -c 1 \
  -f -o gpurun_out/r02_k1_256 python tools/sweep_k1.py 3 0c > gpurun_out/r2b_ncu_k1.log 2>&1
echo "ncu k1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'volume_agg_rowgroup' -s 3 -c 1 \
  -f -o gpurun_out/r02_k1_256_nv5 python tools/sweep_k1.py 5 0c > gpurun_out/r2b_ncu_k1_nv5.log 2>&1
echo "ncu k1 nv5 rc=$?"
ls -la gpurun_out/*.ncu-rep
