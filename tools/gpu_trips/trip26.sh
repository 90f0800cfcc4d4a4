#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stem_tc" -s 1 -c 1 -o gpurun_out/prof_r01_stem -f python bench.py --steps 2 --warmup 1 --cpu-images 0 > gpurun_out/ncu_stem.log 2>&1
ls -la gpurun_out/prof_r01_stem.ncu-rep
