#!/bin/bash
mkdir -p gpurun_out
python tools/debug_value_paths.py > gpurun_out/r2f_debug_value.txt 2>&1; cat gpurun_out/r2f_debug_value.txt | tail -12
python -m pytest tests -m gpu -q -rP --durations=5 > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
grep -E "passed|failed|^FAILED" gpurun_out/r2f_pytest.log | tail -8
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'sdf_mlp|blend_kernel' -s 6 -c 10 -f -o gpurun_out/r02_render python tools/time_render.py 32768 32768 > gpurun_out/r2f_ncu_render.log 2>&1
echo "ncu render rc=$?"; tail -3 gpurun_out/r2f_ncu_render.log
python tools/time_render.py 65536 65536 > gpurun_out/r2f_time_render.txt 2>&1; tail -5 gpurun_out/r2f_time_render.txt
