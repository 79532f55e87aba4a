#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2l_launches_render.csv \
  python tools/time_render.py 65536 65536 > gpurun_out/r2l_render_under_ncu.txt 2>&1
echo "rc=$?"
