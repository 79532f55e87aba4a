#!/bin/bash
# GPU-box visit: full GPU suite, both bench arms, K1 sweep of the shipped variants.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/sweep_k1.py 3 10,25,20 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err
echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench6.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'roofline', d['roofline']['frac'], d['roofline']['ms'], 'e2e ms', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])
r=d['render']; print({k:r[k] for k in ('value','ms_per_step','gpu_launches_per_step')})
PY
