#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/ubench/epi_ubench > gpurun_out/ubench.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 400 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_main.log 2>&1
( time timeout 280 python bench.py --impl reference ) > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log
nproc >> gpurun_out/bench_ref.log
cat gpurun_out/ubench.log
grep -E "passed|failed" gpurun_out/pytest.log | tail -2; grep -E "^FAILED" gpurun_out/pytest.log | head -30
for f in bench_main bench_ref; do echo == $f; tail -n 8 gpurun_out/$f.log | cut -c1-600; grep -o '"kernel_ms": {[^}]*}' gpurun_out/$f.log; done
