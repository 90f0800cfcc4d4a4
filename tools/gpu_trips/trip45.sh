#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed" gpurun_out/pytest.log | tail -2; grep -E "^FAILED" gpurun_out/pytest.log | head
for w in resnet50_int8 resnet50_uint8; do timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --cpu-images 0 > gpurun_out/bench_$w.log 2>&1; tail -1 gpurun_out/bench_$w.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['whole_graph']['kernel_ms'])"; done
timeout 600 python bench.py --workload yolov3_tiny_uint8 --batch 128 --steps 10 --warmup 3 --cpu-images 0 > gpurun_out/bench_yolo128.log 2>&1; tail -1 gpurun_out/bench_yolo128.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('yolo128', d['value'], d['ms_per_step'], d['whole_graph']['kernel_ms'])"
