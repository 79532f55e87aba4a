#!/bin/bash
# ncu --set full of K1 at 256^3 for the variants given as $1 (comma list); one report per variant.
mkdir -p gpurun_out
for v in ${1//,/ }; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'volume_agg_(packed|rowgroup)' -s 3 -c 1 \
    -f -o gpurun_out/k1_v$v python tools/sweep_k1.py 3 $v > gpurun_out/ncu_k1_v$v.log 2>&1
  echo "ncu variant $v exit $?"
done
ls -la gpurun_out/*.ncu-rep
