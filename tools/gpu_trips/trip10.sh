#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest.log
if [ $rc -ne 0 ]; then
  TB200_GEMM_STORE_CS=1 timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_cs1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cs1.log
  grep -E "passed|failed" gpurun_out/pytest_cs1.log | tail -2; grep -E "^FAILED" gpurun_out/pytest_cs1.log | head -40
fi
timeout 400 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_main.log 2>&1
grep -E "passed|failed" gpurun_out/pytest.log | tail -2; grep -E "^FAILED" gpurun_out/pytest.log | head -60
for f in bench_main; do echo == $f; tail -n 3 gpurun_out/$f.log | cut -c1-300; grep -o '"kernel_ms": {[^}]*}' gpurun_out/$f.log; done
