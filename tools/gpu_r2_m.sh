#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
grep -E "passed|failed|^FAILED" gpurun_out/r2m_pytest.log | tail -8
python tools/time_render.py 65536 65536 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench.json'))
print('build', d['ms_per_step'], 'k1', d['roofline']['ms'], d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
r=d['render']; print('render', r['ms_per_step'], r['value'], r['algorithmic']['mlp_tensor_frac_of_tf32_peak'], r['reference_ops_on_gpu']['value'], r['reference_ops_on_gpu'].get('parity_vs_reference_max_rel'))
print('lattice', d['lattice']['ms_per_step'], 'train', d['train_step']['ms_per_step'])
PY
