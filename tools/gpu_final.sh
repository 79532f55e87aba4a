#!/bin/bash
# End-of-round GPU visit: suite, both bench arms, launch list of the bench command, ncu --set full of K1.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
echo "bench reference exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --render-steps 1 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
echo "ncu launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:volume_agg_rowgroup -s 3 -c 1 \
  -f -o gpurun_out/r01_k1_256 python tools/sweep_k1.py 3 0 > gpurun_out/ncu_k1.log 2>&1
echo "ncu k1 exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'roofline', d['roofline']['frac'], d['roofline']['ms'], 'e2e ms', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])
r=d['render']; print({k:r[k] for k in ('value','ms_per_step','gpu_launches_per_step')})
print(open('gpurun_out/bench_final_reference.json').read()[:600])
PY
