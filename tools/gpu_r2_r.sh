#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_volume_gpu.py -m gpu -q -x -k "two_gpus" 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --no-render --no-lattice --no-train --no-cpu > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2r_bench_n2.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(json.dumps(d['regularise'],indent=1)); print(d['ms_per_step'])
PY
tail -5 gpurun_out/r2r_bench_n2.err
