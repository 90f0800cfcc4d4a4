#!/bin/bash
# round 2, final evidence on one B200: default bench line (+ reference arm), the other BASELINE.json workloads, launch list of the
# default command, ncu --set full of one whole-batch step (never a bench value), full GPU test log.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r02_final_gpu.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_final_bench_default.log 2>&1; tail -n 1 gpurun_out/r02_final_bench_default.log | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference_arm.log 2>&1; tail -n 1 gpurun_out/r02_final_bench_reference_arm.log | cut -c1-600
for w in mobilenet_v1_uint8 resnet50_int8 resnet50_uint8 yolov3_tiny_uint8:128 yolov3_tiny_uint8:16 yolov3_tiny_int8:128 yolov5s_int8:64 yolov5s_uint8:64 yolov5s_int8:8; do
  n=${w%%:*}; b=0; [ "$n" != "$w" ] && b=${w##*:}
  timeout 300 python bench.py --workload $n --batch $b --steps 20 --warmup 3 --cpu-window 0 > gpurun_out/r02_final_bench_${n}_b$b.log 2>&1
  tail -n 1 gpurun_out/r02_final_bench_${n}_b$b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['roofline']['bound'], round(d['roofline']['frac'],3), d['roofline']['kernel'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --cpu-window 0 > gpurun_out/r02_final_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_i8|conv_dw3x3|stem_tc|pool_kernel|nhwc_to_nchw" -s 30 -c 30 -o gpurun_out/prof_r02_final -f python bench.py --steps 2 --warmup 1 --cpu-window 0 > gpurun_out/r02_final_ncu_run.log 2>&1
ls -la gpurun_out/prof_r02_final.ncu-rep
timeout 300 python tools/layer_times.py 256 mobilenet_v1 int8 > gpurun_out/r02_final_layers_mobilenet_b256.txt 2>&1; tail -n 1 gpurun_out/r02_final_layers_mobilenet_b256.txt
timeout 300 python tools/layer_times.py 512 resnet50 int8 > gpurun_out/r02_final_layers_resnet50_i8_b512.txt 2>&1; tail -n 1 gpurun_out/r02_final_layers_resnet50_i8_b512.txt
timeout 300 python tools/layer_times.py 512 resnet50 uint8 > gpurun_out/r02_final_layers_resnet50_u8_b512.txt 2>&1; tail -n 1 gpurun_out/r02_final_layers_resnet50_u8_b512.txt
timeout 300 python tools/layer_times.py 128 yolov3_tiny uint8 > gpurun_out/r02_final_layers_yolov3_tiny_u8_b128.txt 2>&1; tail -n 1 gpurun_out/r02_final_layers_yolov3_tiny_u8_b128.txt
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_pytest_gpu.log
grep -E "passed|failed" gpurun_out/r02_final_pytest_gpu.log | tail -2
