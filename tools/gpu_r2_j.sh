#!/bin/bash
# round 2, visit J: reverse kernel with overlapped feature-part MMAs -> GPU suite, render timing; sanitizer on the new kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
grep -E "passed|failed|^FAILED" gpurun_out/r2j_pytest.log | tail -8
python tools/time_render.py 65536 65536 > gpurun_out/r2j_time_render.txt 2>&1; tail -2 gpurun_out/r2j_time_render.txt
python tools/prof_mlp_tc.py > gpurun_out/r2j_prof_mlp.txt 2>&1; tail -12 gpurun_out/r2j_prof_mlp.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_lncc_gpu.py tests/test_marching_cubes_gpu.py \
  "tests/test_volume_gpu.py::test_rowgroup_kernel_culling_is_bit_identical[64-hw0-3]" -q -x > gpurun_out/r2j_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2j_memcheck.log | tail -4
