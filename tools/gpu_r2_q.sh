#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reg_network_gpu.py -m gpu -q -x 2>&1 | tail -4
