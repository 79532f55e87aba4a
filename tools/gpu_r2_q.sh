#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_blend.py > gpurun_out/r2q_blend.txt 2>&1; grep -v Warn gpurun_out/r2q_blend.txt | tail -8
