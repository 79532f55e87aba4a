#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench5.json 2> gpurun_out/bench5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench5.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['e2e']['ms_per_step'], d['config']['wall_s_timed_loop'])
r=d['render']; print({k:r[k] for k in ('value','ms_per_step','gpu_launches_per_step')})
PY
timeout 300 python tools/profile_render.py 16384 60 > gpurun_out/profile_render.txt 2>&1
tail -70 gpurun_out/profile_render.txt | cut -c1-200
