#!/bin/bash
# GPU-box visit: full GPU test-suite (no -x), ncu --set full captures of K1 and of the render kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'volume_agg_(rowgroup|packed)' -s 3 -c 1 \
  -f -o gpurun_out/r01_k1_256 python tools/sweep_k1.py 3 0 > gpurun_out/ncu_k1.log 2>&1
echo "ncu k1 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'sdf_mlp|trilinear|lookup_feature|upsample_kernel|merge_kernel|mask_nearest|encode_kernel|decode_kernel|tv_reduce' \
  -s 30 -c 30 -f -o gpurun_out/r01_render python tools/time_render.py 16384 16384 > gpurun_out/ncu_render.log 2>&1
echo "ncu render exit $?"
ls -la gpurun_out/*.ncu-rep
