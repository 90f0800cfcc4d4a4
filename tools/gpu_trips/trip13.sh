#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_i8" -s 14 -c 14 -o gpurun_out/prof_r01_t13 -f python bench.py --steps 2 --warmup 1 --cpu-images 0 > gpurun_out/ncu_t13.log 2>&1
ls -la gpurun_out/prof_r01_t13.ncu-rep
