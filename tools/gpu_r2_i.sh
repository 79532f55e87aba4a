#!/bin/bash
# round 2, final multi-GPU visit (8 GPUs): fused slab exchange at 8 ranks, bench at N = 8, 4, 2
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  tools/check_fused_slabs.py > gpurun_out/r2i_fused8.log 2>&1
echo "check_fused8 rc=$?"; grep -c "bit-identical to the 1-GPU build: True" gpurun_out/r2i_fused8.log; grep "world" gpurun_out/r2i_fused8.log
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n \
  bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2i_bench_n$n.json 2> gpurun_out/r2i_bench_n$n.err
echo "bench n$n rc=$?"
done
python - <<'PY'
import json
for n in (8,4,2):
    d=json.load(open(f'gpurun_out/r2i_bench_n{n}.json'))
    print(n, 'build ms', round(d['ms_per_step'],4), d['verified']['slabs_bit_identical'], 'nvlink', round(d['roofline']['achieved']), 'render ms', round(d['render']['ms_per_step'],1), d['render']['verified']['ray_shards_bit_identical'], 'lattice ms', round(d['lattice']['ms_per_step'],1), d['lattice']['verified']['lattice_slabs_bit_identical'], 'train', round(d['train_step']['ms_per_step'],1))
PY
