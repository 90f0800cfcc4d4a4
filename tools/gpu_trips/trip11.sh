#!/bin/bash
mkdir -p gpurun_out
# one whole-batch step = 30 kernels; skip prerun/warmup launches (warmup 1 -> 30 kernels), profile the next 30
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_i8|conv_dw3x3|conv_stem|pool|nhwc" -s 30 -c 30 -o gpurun_out/prof_r01_t11 -f python bench.py --steps 2 --warmup 1 --cpu-images 0 > gpurun_out/ncu_t11.log 2>&1
tail -3 gpurun_out/ncu_t11.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
