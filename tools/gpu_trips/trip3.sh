#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_tc.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-images 0 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_i8|conv_dw3x3|conv_stem" -s 28 -c 28 -o gpurun_out/prof_all -f python bench.py --steps 2 --warmup 1 --cpu-images 0 > gpurun_out/ncu_all.log 2>&1
tail -n 4 gpurun_out/pytest.log; tail -n 1 gpurun_out/bench_tc.log
