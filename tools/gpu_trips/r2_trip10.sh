#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/layer_times.py 512 resnet50 uint8 2>&1 | grep -E "layer  [2-9] |layer 1[0-5] |total"
echo "--- without the constant-tile MMA (wrong results, timing only)"
TB200_DEBUG_NO_CPLANE=1 timeout 300 python tools/layer_times.py 512 resnet50 uint8 2>&1 | grep -E "layer  [2-9] |layer 1[0-5] |total"
echo "--- int8"
timeout 300 python tools/layer_times.py 512 resnet50 int8 2>&1 | grep -E "layer  [2-9] |layer 1[0-5] |total"
# ncu of the window kernels and one uint8 / int8 GEMM pair
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_window" -c 3 -o gpurun_out/prof_r02_window -f python tools/layer_times.py 128 yolov3_tiny uint8 > gpurun_out/ncu_window.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_i8" -s 1 -c 3 -o gpurun_out/prof_r02_gemm_u8 -f python tools/layer_times.py 512 resnet50 uint8 > gpurun_out/ncu_gemm_u8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_i8" -s 1 -c 3 -o gpurun_out/prof_r02_gemm_i8 -f python tools/layer_times.py 512 resnet50 int8 > gpurun_out/ncu_gemm_i8.log 2>&1
ls -la gpurun_out/*.ncu-rep
