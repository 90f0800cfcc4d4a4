#!/bin/bash
mkdir -p gpurun_out
for w in resnet50_uint8 resnet50_int8 yolov3_tiny_uint8 mobilenet_v1_uint8; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --cpu-images 0 > gpurun_out/bench_$w.log 2>&1
  echo == $w; tail -1 gpurun_out/bench_$w.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['whole_graph']['kernel_ms'], d['whole_graph']['achieved_tops'])
except Exception as ex: print('ERR', ex)
"
done
timeout 600 python bench.py --workload yolov3_tiny_uint8 --batch 128 --steps 10 --warmup 3 --cpu-images 0 > gpurun_out/bench_yolo128.log 2>&1; tail -1 gpurun_out/bench_yolo128.log | cut -c1-200
