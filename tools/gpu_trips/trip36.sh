#!/bin/bash
mkdir -p gpurun_out
# launch list (durations only) of two whole-batch steps, then one --set full capture of the second step's 30 kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_i8|conv_dw3x3|stem_tc|pool|nhwc" -c 90 --csv --log-file gpurun_out/r01_final_launches.csv python bench.py --steps 2 --warmup 1 --cpu-images 0 > gpurun_out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_i8|conv_dw3x3|stem_tc|pool|nhwc" -s 30 -c 30 -o gpurun_out/prof_r01_final -f python bench.py --steps 2 --warmup 1 --cpu-images 0 > gpurun_out/ncu_final.log 2>&1
ls -la gpurun_out/prof_r01_final.ncu-rep gpurun_out/r01_final_launches.csv
timeout 300 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-300
