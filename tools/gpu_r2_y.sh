#!/bin/bash
# round 2, visit Y (2 GPUs): final tree -> whole GPU suite incl. the 2-GPU tests, smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2y_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
