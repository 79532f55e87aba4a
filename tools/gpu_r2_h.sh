#!/bin/bash
# round 2, visit H: final K1 -> GPU suite, ncu --set full capture of the shipped K1 (traffic), launch list of the bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rP --durations=8 > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
grep -E "passed|failed|^FAILED" gpurun_out/r2h_pytest.log | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'volume_agg_rowgroup' -s 3 -c 1 \
  -f -o gpurun_out/r02_k1_256 python tools/sweep_k1.py 3 0c > gpurun_out/r2h_ncu_k1.log 2>&1
echo "ncu k1 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --steps 4 --warmup 3 --render-steps 1 --no-cpu > gpurun_out/r2h_bench_under_ncu.json 2> gpurun_out/r2h_bench_under_ncu.err
echo "ncu launches rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r2h_bench.json
