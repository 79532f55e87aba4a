#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_mlp_tc_gpu.py tests/test_render_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/prof_rev_phases.py > gpurun_out/r2q_rev_phases.txt 2>&1; grep -v Warn gpurun_out/r2q_rev_phases.txt | tail -24
timeout 300 python tools/time_render.py 65536 65536 2>&1 | tail -1
