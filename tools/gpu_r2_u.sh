#!/bin/bash
# round 2, visit U (1 GPU): final tree -> GPU suite, bench N=1 (both arms), launch list, ncu of render kernels + K13
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rP --durations=8 > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2u_pytest.log | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2u_bench_ref.json 2> gpurun_out/r2u_bench_ref.err
echo "bench ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02b_launches_bench.csv \
  python bench.py --steps 4 --warmup 3 --render-steps 1 --no-cpu --no-lattice --no-train > gpurun_out/r2u_bench_under_ncu.json 2> gpurun_out/r2u_bench_under_ncu.err
echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'sdf_mlp|blend_kernel' -s 6 -c 10 -f -o gpurun_out/r02b_render python tools/time_render.py 32768 32768 > gpurun_out/r2u_ncu_render.log 2>&1
echo "ncu render rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'conv3d_k3|deconv3d|norm_relu' -s 40 -c 20 -f -o gpurun_out/r02b_regnet python tools/prof_regnet.py > gpurun_out/r2u_ncu_regnet.log 2>&1
echo "ncu regnet rc=$?"
python tools/prof_rev_phases.py > gpurun_out/r2u_phases.txt 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2u_bench.json') if l.startswith('{')][-1])
print('build', d['ms_per_step'], 'k1', d['roofline']['ms'], d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'e2e', d['e2e']['ms_per_step'])
r=d['render']; print('render', r['ms_per_step'], r['value'], r['algorithmic']['mlp_tensor_frac_of_tf32_peak'], r['reference_ops_on_gpu']['value'])
print('lattice', d['lattice']['ms_per_step'], 'train', d['train_step']['ms_per_step'], d['train_step']['reference_ops_on_gpu']['ms_per_step'])
g=d['regularise']; print('regularise', g['ms_per_step'], g['network_only_ms'], g['reference_ops_on_gpu'])
PY
