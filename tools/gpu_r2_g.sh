#!/bin/bash
mkdir -p gpurun_out
python tools/debug_value_paths.py > gpurun_out/r2g_debug_value.txt 2>&1; grep "^D" gpurun_out/r2g_debug_value.txt
python -m pytest tests -m gpu -q -rP --durations=8 > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
grep -E "passed|failed|^FAILED" gpurun_out/r2g_pytest.log | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r2g_bench.json; tail -3 gpurun_out/r2g_bench.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err
echo "bench ref rc=$?"; cut -c1-400 gpurun_out/r2g_bench_ref.json
