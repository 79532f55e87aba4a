#!/bin/bash
mkdir -p gpurun_out
python tools/time_blend.py 2>&1 | grep -v Warn | tail -6
python -m pytest tests/test_reference_render_gpu.py tests/test_mlp_tc_gpu.py -m gpu -q -rP 2>&1 | grep -E "passed|failed|share beyond|Error" | head -12
