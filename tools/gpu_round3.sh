#!/bin/bash
mkdir -p gpurun_out
./tools/_bin/ubench_store > gpurun_out/ubench_store.txt 2>&1
cat gpurun_out/ubench_store.txt
timeout 300 ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_lg.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,sm__cycles_active.avg,gpu__time_duration.sum \
  --clock-control none --csv --log-file gpurun_out/ubench_store_ncu.csv ./tools/_bin/ubench_store > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
echo skip bench

