#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gemm_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/probe.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_tc.log 2>&1
# launch list of one bench run (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-images 0 > gpurun_out/ncu_bench.log 2>&1
# full capture of the dominant kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_i8 -s 14 -c 6 -o gpurun_out/prof_gemm -f python bench.py --steps 2 --warmup 1 --cpu-images 0 > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_dw3x3 -s 13 -c 4 -o gpurun_out/prof_dw -f python bench.py --steps 2 --warmup 1 --cpu-images 0 > gpurun_out/ncu_dw.log 2>&1
tail -3 gpurun_out/probe.log gpurun_out/pytest.log; tail -1 gpurun_out/bench_tc.log
