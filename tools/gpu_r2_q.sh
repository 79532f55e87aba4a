#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_final_pytest.log
