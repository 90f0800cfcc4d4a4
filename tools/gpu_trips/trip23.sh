#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_base.log 2>&1
TB200_GEMM_PAIR=1 timeout 400 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_pair.log 2>&1
for f in bench_base bench_pair; do echo == $f; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1; grep -o '"kernel_ms": {[^}]*}' gpurun_out/$f.log; done
