#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29591 tools/check_peer_regulariser.py > gpurun_out/r2x_peer_$1.log 2>&1; echo "rc=$?"
grep -E "^rank|^world" gpurun_out/r2x_peer_$1.log | sort | cut -c1-200
tail -3 gpurun_out/r2x_peer_$1.log | cut -c1-300
