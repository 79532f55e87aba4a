#!/bin/bash
# round 2, visit S: checkpoint after K10 pairs / K4 reverse reorder / regulariser hand-off + K13: GPU suite + bench N=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rP --durations=8 > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
grep -E "passed|failed|^FAILED" gpurun_out/r2s_pytest.log | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2s_bench.json') if l.startswith('{')][-1])
print('build', d['ms_per_step'], 'k1', d['roofline']['ms'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])
r=d['render']; print('render', r['ms_per_step'], r['value'], r['algorithmic']['mlp_tensor_frac_of_tf32_peak'], r['reference_ops_on_gpu']['value'])
print('lattice', d['lattice']['ms_per_step'], 'train', d['train_step']['ms_per_step'], d['train_step'].get('reference_ops_on_gpu'))
print('regularise', json.dumps(d['regularise'])[:1500])
PY
tail -3 gpurun_out/r2s_bench.err
