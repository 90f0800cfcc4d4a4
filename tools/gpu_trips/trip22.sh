#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_tma.log 2>&1
TB200_GEMM_DIRECT=1 timeout 400 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_direct.log 2>&1
TB200_GEMM_DIRECT=1 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -k "gemm or mobilenet or fixture or tiny" > gpurun_out/pytest_direct.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_direct.log
tail -3 gpurun_out/pytest_direct.log
for f in bench_tma bench_direct; do echo == $f; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1; grep -o '"kernel_ms": {[^}]*}' gpurun_out/$f.log; done
