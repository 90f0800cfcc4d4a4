#!/bin/bash
mkdir -p gpurun_out
TB200_GEMM_TEAMS=1 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -k "gemm or mobilenet or conv or tiny" > gpurun_out/pytest_teams.log 2>&1; tail -2 gpurun_out/pytest_teams.log
timeout 400 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_base.log 2>&1
TB200_GEMM_TEAMS=1 timeout 400 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_teams.log 2>&1
for f in bench_base bench_teams; do echo == $f; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1; grep -o '"kernel_ms": {[^}]*}' gpurun_out/$f.log; done
