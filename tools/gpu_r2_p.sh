#!/bin/bash
# round 2, visit P: final tree -> GPU suite, bench N=1 (both arms), ncu of the current render kernels, phase timers
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rP --durations=8 > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
grep -E "passed|failed|^FAILED" gpurun_out/r2p_pytest.log | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2p_bench_ref.json 2> gpurun_out/r2p_bench_ref.err
echo "bench ref rc=$?"
python tools/prof_rev_phases.py > gpurun_out/r2p_phases.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'sdf_mlp|blend_kernel' -s 6 -c 10 -f -o gpurun_out/r02_render python tools/time_render.py 32768 32768 > gpurun_out/r2p_ncu_render.log 2>&1
echo "ncu render rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_bench.json'))
print('build', d['ms_per_step'], 'k1', d['roofline']['ms'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])
r=d['render']; print('render', r['ms_per_step'], r['value'], r['algorithmic']['mlp_tensor_frac_of_tf32_peak'], r['reference_ops_on_gpu']['value'])
print('lattice', d['lattice']['ms_per_step'], 'train', d['train_step']['ms_per_step'])
PY
