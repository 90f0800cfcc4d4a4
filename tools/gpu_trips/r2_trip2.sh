#!/bin/bash
# round 2, trip 2: engine refactor (multi-GPU contexts, host registration, arena reuse, new glue ops, tmfile integration)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "passed|failed" gpurun_out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/pytest.log | head -40; tail -40 gpurun_out/pytest.log | cut -c1-400
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_main.log 2>&1; tail -n 2 gpurun_out/bench_main.log | cut -c1-3000
timeout 300 python bench.py --steps 20 --warmup 3 --cpu-window 0 --pinned > gpurun_out/bench_pinned.log 2>&1; tail -n 1 gpurun_out/bench_pinned.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pinned e2e', d['e2e']['value'], d['value'])"
