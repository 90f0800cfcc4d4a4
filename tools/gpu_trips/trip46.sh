#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed" gpurun_out/pytest.log | tail -2; grep -E "^FAILED" gpurun_out/pytest.log | head
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-200
for w in resnet50_uint8 yolov3_tiny_uint8; do timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --cpu-images 0 > gpurun_out/bench_$w.log 2>&1; tail -1 gpurun_out/bench_$w.log | cut -c1-160; done
timeout 600 python bench.py --workload yolov3_tiny_uint8 --batch 128 --steps 10 --warmup 3 --cpu-images 0 > gpurun_out/bench_yolo128.log 2>&1; tail -1 gpurun_out/bench_yolo128.log | cut -c1-160
