#!/bin/bash
mkdir -p gpurun_out
TB200_DW_NO_PACK3=1 timeout 400 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_nopack3.log 2>&1
TB200_PIPELINE_CHUNKS=4 timeout 400 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_chunks4.log 2>&1
TB200_PIPELINE_CHUNKS=8 timeout 400 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_chunks8.log 2>&1
for f in bench_nopack3 bench_chunks4 bench_chunks8; do echo == $f; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -2 | tr '\n' ' '; grep -o '"kernel_ms": {[^}]*}' gpurun_out/$f.log; done
