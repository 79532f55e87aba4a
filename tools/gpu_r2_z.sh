#!/bin/bash
# round 2, visit Z: compute-sanitizer memcheck over the kernels added late in the round (K13, K10 on FFMA2 pairs)
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_reg_network_gpu.py -m gpu -q -x -k "k13 or instnorm or slab_ops" > gpurun_out/r2z_memcheck_k13.log 2>&1; echo "memcheck k13 rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2z_memcheck_k13.log | tail -3
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_render_gpu.py -m gpu -q -x -k "k10" > gpurun_out/r2z_memcheck_k10.log 2>&1; echo "memcheck k10 rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2z_memcheck_k10.log | tail -3
