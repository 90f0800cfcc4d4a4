#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --section SchedulerStats --section WarpStateStats --clock-control none -k regex:"gemm_i8|conv_gather|conv_dw|stem|pool|pointwise|relu_same|concat|upsample|conv_direct|nhwc|nchw" -s 48 -c 48 -o gpurun_out/prof_r01_yolo -f python bench.py --workload yolov3_tiny_uint8 --batch 128 --steps 2 --warmup 1 --cpu-images 0 > gpurun_out/ncu_yolo.log 2>&1
ls -la gpurun_out/prof_r01_yolo.ncu-rep; tail -2 gpurun_out/ncu_yolo.log | cut -c1-200
