#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reg_network_gpu.py -m gpu -q -x -s 2>&1 | tail -12
timeout 300 python tools/prof_regnet.py > gpurun_out/r2q_regnet3.txt 2>&1; grep -v Warn gpurun_out/r2q_regnet3.txt | head -34 | cut -c1-90,180-260
