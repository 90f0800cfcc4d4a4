#!/bin/bash
# round 2, trip 19: rare-path queue, "push the whole guarded word, re-derive in the drain" (build/alt, -DTB200_FIXQ_WORD) vs the product
mkdir -p gpurun_out
export TB200_LIB=$PWD/build/alt/libtengine_b200_fixqword.so
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 300 -x > gpurun_out/pytest_alt.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_alt.log
grep -E "passed|failed" gpurun_out/pytest_alt.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/pytest_alt.log | head
for lib in alt product; do
  if [ $lib = alt ]; then export TB200_LIB=$PWD/build/alt/libtengine_b200_fixqword.so; else unset TB200_LIB; fi
  for w in mobilenet_v1_int8 resnet50_int8 resnet50_uint8; do
    timeout -k 10 300 python bench.py --workload $w --steps 20 --warmup 3 --cpu-window 0 > gpurun_out/bench_${lib}_$w.log 2>&1
    tail -n 1 gpurun_out/bench_${lib}_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', d['config']['workload'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['whole_graph']['kernel_ms_gpu0'])"
  done
done
