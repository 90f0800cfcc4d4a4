#!/bin/bash
# round 2, trip 4: 8-column uint8 epilogue, pipelined window kernel, YOLOv5s, pack cache; debug of the ResNet tmfile case; ncu of C4
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed" gpurun_out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/pytest.log | head -40
timeout 300 python tools/debug_resnet_b4.py resnet50 > gpurun_out/debug_resnet_b4.log 2>&1; tail -20 gpurun_out/debug_resnet_b4.log
timeout 300 python tools/layer_times.py 128 yolov3_tiny uint8 > gpurun_out/layers_yolo_u8_b128.txt 2>&1; tail -40 gpurun_out/layers_yolo_u8_b128.txt
timeout 400 python tools/layer_times.py 512 resnet50 uint8 > gpurun_out/layers_resnet50_u8_b512.txt 2>&1; tail -2 gpurun_out/layers_resnet50_u8_b512.txt
timeout 300 python tools/layer_times.py 8 yolov5s int8 > gpurun_out/layers_yolov5s_i8_b8.txt 2>&1; tail -2 gpurun_out/layers_yolov5s_i8_b8.txt
for w in yolov3_tiny_uint8 resnet50_uint8; do
  b=0; [ $w = yolov3_tiny_uint8 ] && b=128
  timeout 300 python bench.py --workload $w --batch $b --steps 10 --warmup 3 --cpu-window 0 > gpurun_out/bench_$w.log 2>&1
  tail -n 1 gpurun_out/bench_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['whole_graph']['kernel_ms_gpu0'])"
done
# ncu: one whole-batch step of YOLOv3-tiny uint8 b=128 without CUDA graph (layer_times runs 3+5 profile passes of 35+ launches; skip the first pass)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_window|gemm_i8|concat|pointwise|pool" -s 40 -c 36 -o gpurun_out/prof_r02_yolo_u8_b128 -f python tools/layer_times.py 128 yolov3_tiny uint8 > gpurun_out/ncu_yolo.log 2>&1
ls -la gpurun_out/*.ncu-rep
