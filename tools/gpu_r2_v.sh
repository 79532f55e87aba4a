#!/bin/bash
# round 2, visits V (N = 2 / 4 / 8 GPUs): multi-GPU tests (N = 2 only) and the full bench line at N ranks
N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  python -m pytest tests -m gpu -q -k "two_gpus" > gpurun_out/r2v_pytest2.log 2>&1; echo "pytest2 rc=$?"; tail -2 gpurun_out/r2v_pytest2.log
fi
if [ "$N" = "8" ] && [ -z "$SKIP_CHECK" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/check_fused_slabs.py > gpurun_out/r2v_fused8.log 2>&1; echo "check rc=$?"
  grep -E "world 8" gpurun_out/r2v_fused8.log; grep -c "bit-identical to the 1-GPU build: True" gpurun_out/r2v_fused8.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2v_bench_n$N.json 2> gpurun_out/r2v_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2v_bench_n$N.json') if l.startswith('{')][-1])
print('build', d['ms_per_step'], 'nvlink', d['roofline']['frac'], d['verified']['slabs_bit_identical'])
r=d['render']; print('render', r['ms_per_step'], r.get('verified'))
print('lattice', d['lattice']['ms_per_step'], d['lattice'].get('verified'), 'train', d['train_step']['ms_per_step'])
g=d['regularise']; print('regularise', g['ms_per_step'], 'nccl', g.get('nccl_eager_ms'), 'peer', g.get('peer_memory_ms'), 'one gpu', g['one_gpu_ms'], g.get('peer_memory'), g['verified']['as_accurate_as_the_whole_volume_pipeline'])
PY
tail -2 gpurun_out/r2v_bench_n$N.err
