#!/bin/bash
# round 2, trip 8: tmfile debug (constants), uneven pipeline chunks for the e2e path, int8 MMA peak probe
mkdir -p gpurun_out
timeout 400 python tools/debug_tmfile.py > gpurun_out/debug_tmfile.log 2>&1; grep -B1 -A2 "consts" gpurun_out/debug_tmfile.log | head -30 | cut -c1-400
python - <<'PY'
import sys
sys.path.insert(0, '.')
from tengine_b200 import runtime as rt
c = rt.Context(0)
print("int8 MMA peak probe: %.1f TOP/s" % c.probe_int8_tops())
c.close()
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 600 -k "pipelined or conv_kernels or gemm_1x1" > gpurun_out/pytest_part.log 2>&1; tail -3 gpurun_out/pytest_part.log
for split in "" "64,192" "40,88,128" "32,96,128" "48,208"; do
  TB200_PIPELINE_SPLIT=$split timeout 300 python bench.py --steps 20 --warmup 3 --cpu-window 0 > gpurun_out/bench_split.log 2>&1
  tail -n 1 gpurun_out/bench_split.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('split [$split]', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['roofline']['bound'], round(d['roofline']['frac'],3))"
done
