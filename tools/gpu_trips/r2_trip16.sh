#!/bin/bash
# round 2, trips 16-17: GEMM rare path deferred to a per-CTA queue (drained after the last tile) vs inline (TB200_NO_FIXQ=1)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed" gpurun_out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/pytest.log | head -40
for q in fixq inline; do
  if [ $q = inline ]; then export TB200_NO_FIXQ=1; else unset TB200_NO_FIXQ; fi
  for w in mobilenet_v1_int8 resnet50_int8 resnet50_uint8 yolov3_tiny_uint8 yolov5s_int8; do
    b=0; [ $w = yolov3_tiny_uint8 ] && b=128; [ $w = yolov5s_int8 ] && b=64
    timeout 300 python bench.py --workload $w --batch $b --steps 20 --warmup 3 --cpu-window 0 > gpurun_out/bench_${q}_$w.log 2>&1
    tail -n 1 gpurun_out/bench_${q}_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$q', d['config']['workload'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['whole_graph']['kernel_ms_gpu0'])"
  done
done
