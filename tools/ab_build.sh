#!/bin/bash
# A/B of K1 tuning-knob values through bench.py's own build timing: bash tools/ab_build.sh "0 30 31 32"
for v in $1; do
python - <<PY | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v: build', round(d['ms_per_step']*1e3,1), 'us, K1@256', round(d['roofline']['ms']*1e3,1), 'us')"
import sys
sys.argv=["bench.py","--no-cpu","--no-render","--steps","60","--warmup","5"]
from gens_b200 import _lib
_lib.lib().gens_debug_set_variant($v)
import bench
bench.main()
PY
done
