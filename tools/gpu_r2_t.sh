#!/bin/bash
# round 2, visit T (8 GPUs): peer-store order experiment + lean bench (build + regularise hand-off)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/check_fused_slabs.py > gpurun_out/r2t_fused8.log 2>&1; echo "check rc=$?"
grep -E "bit-identical|world 8" gpurun_out/r2t_fused8.log | sort | uniq -c
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 8 --steps 20 --warmup 5 --no-render --no-lattice --no-train --no-cpu > gpurun_out/r2t_bench_n8.json 2> gpurun_out/r2t_bench_n8.err; echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2t_bench_n8.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('build ms', d['ms_per_step'], 'roofline', d['roofline']['frac'], d['verified']); print(json.dumps(d['regularise'])[:1200])
PY
tail -3 gpurun_out/r2t_bench_n8.err
