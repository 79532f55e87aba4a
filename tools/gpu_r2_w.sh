#!/bin/bash
# round 2, visit W (1 GPU): bench N=1 (both arms) to files; small ncu captures summarised ON the box (reports deleted:
# gpurun copies at most 64 MiB back)
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2w_bench_ref.json 2> gpurun_out/r2w_bench_ref.err
echo "bench ref rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:'conv3d_k3|deconv3d|norm_relu' -s 17 -c 17 -f -o /tmp/r02b_regnet python tools/ncu_k13.py > gpurun_out/r2w_ncu_regnet.log 2>&1
echo "ncu regnet rc=$?"
python tools/summarize_ncu.py kernels /tmp/r02b_regnet.ncu-rep gpurun_out/r02_k13_kernels.txt
timeout 300 ncu --set full --clock-control none -k regex:'sdf_mlp|blend_kernel' -s 14 -c 2 -f -o /tmp/r02b_render python tools/time_render.py 32768 32768 > gpurun_out/r2w_ncu_render.log 2>&1
echo "ncu render rc=$?"
python tools/summarize_ncu.py kernels /tmp/r02b_render.ncu-rep gpurun_out/r02b_render_kernels.txt
cat gpurun_out/r02_k13_kernels.txt | cut -c1-200 | tail -20
cat gpurun_out/r02b_render_kernels.txt | cut -c1-200 | tail -4
du -sh gpurun_out
