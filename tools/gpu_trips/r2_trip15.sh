#!/bin/bash
# round 2, trip 15: what does the rare literal path cost?  Same benches with the product library and with the tie guard compiled
# out (tools/build_nofix_lib.sh, timing experiment only).
mkdir -p gpurun_out
for lib in product nofix; do
  for w in mobilenet_v1_int8 resnet50_int8 resnet50_uint8; do
    if [ $lib = nofix ]; then export TB200_LIB=$PWD/build/nofix/libtengine_b200_nofix.so; else unset TB200_LIB; fi
    timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --cpu-window 0 > gpurun_out/bench_${lib}_$w.log 2>&1
    tail -n 1 gpurun_out/bench_${lib}_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', d['config']['workload'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['whole_graph']['kernel_ms_gpu0'])"
  done
done
