#!/bin/bash
# round 2, trip 7: uint8 GEMM with the zero points folded by a second MMA, window kernel (magic division, int8-form epilogue); tmfile debug
mkdir -p gpurun_out
timeout 400 python tools/debug_tmfile.py > gpurun_out/debug_tmfile.log 2>&1; tail -24 gpurun_out/debug_tmfile.log | cut -c1-600
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed" gpurun_out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/pytest.log | head -40
timeout 300 python tools/layer_times.py 128 yolov3_tiny uint8 > gpurun_out/layers_yolo_u8_b128.txt 2>&1; tail -40 gpurun_out/layers_yolo_u8_b128.txt | grep -E "window|igemm|gemm|total"
timeout 400 python tools/layer_times.py 512 resnet50 uint8 > gpurun_out/layers_resnet50_u8_b512.txt 2>&1; tail -2 gpurun_out/layers_resnet50_u8_b512.txt
for w in yolov3_tiny_uint8 resnet50_uint8 mobilenet_v1_int8; do
  b=0; [ $w = yolov3_tiny_uint8 ] && b=128
  timeout 300 python bench.py --workload $w --batch $b --steps 10 --warmup 3 --cpu-window 0 > gpurun_out/bench_$w.log 2>&1
  tail -n 1 gpurun_out/bench_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['whole_graph']['kernel_ms_gpu0'])"
done
