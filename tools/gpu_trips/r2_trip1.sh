#!/bin/bash
# round 2, trip 1: the new parity tests (full-size C3/C4, kernel ABI, dilation) on the round-1 kernels + per-layer times of C3/C4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed" gpurun_out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/pytest.log | head -60
timeout 300 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_main.log 2>&1; tail -n 2 gpurun_out/bench_main.log | cut -c1-600
