#!/bin/bash
# first GPU trip: probe the tcgen05 GEMM, run the GPU test-suite, a short bench (both paths), and a launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python tools/gemm_probe.py > gpurun_out/probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/probe.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-images 40 --no-tensorcore > gpurun_out/bench_cudacore.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-images 0 > gpurun_out/bench_tc.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -5 gpurun_out/probe.log gpurun_out/smoke.log gpurun_out/bench_tc.log gpurun_out/bench_cudacore.log gpurun_out/bench_ref.log
tail -30 gpurun_out/pytest.log
