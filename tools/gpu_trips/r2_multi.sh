#!/bin/bash
# round 2, multi-GPU trip (gpurun --gpus N): NCCL path of the single-process GPU group, bench under torchrun as the driver launches it,
# C4 strong scaling (global batch 128 split over the GPUs), unmodified tm_benchmark on the group
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/multi_gpus.txt 2>&1
# stage-by-stage check first (watchdog dumps the stack and exits if a stage hangs for 60 s)
timeout -k 10 300 python tools/debug_multi.py $N > gpurun_out/debug_multi_${N}gpu.log 2>&1; tail -n 12 gpurun_out/debug_multi_${N}gpu.log | cut -c1-300
if ! grep -q '^OK' gpurun_out/debug_multi_${N}gpu.log; then
  echo 'multi-GPU path did not complete: retrying the stage check with the cudaMemcpyPeer broadcast (TB200_NO_NCCL=1), then stopping'
  TB200_NO_NCCL=1 timeout -k 10 300 python tools/debug_multi.py $N > gpurun_out/debug_multi_${N}gpu_nonccl.log 2>&1; tail -n 12 gpurun_out/debug_multi_${N}gpu_nonccl.log | cut -c1-300
  exit 1
fi
if [ "$2" != notest ]; then
timeout -k 10 420 python -m pytest tests/test_gpu_multi.py tests/test_tengine_integration.py -m gpu -q -p no:cacheprovider --timeout -k 10 200 > gpurun_out/pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_${N}gpu.log
grep -E "passed|failed|skipped" gpurun_out/pytest_multi_${N}gpu.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/pytest_multi_${N}gpu.log | head
fi
NCCL_DEBUG=WARN timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --cpu-window 0 > gpurun_out/bench_mobilenet_${N}gpu.log 2>&1
tail -n 1 gpurun_out/bench_mobilenet_${N}gpu.log | cut -c1-1200
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --workload yolov3_tiny_uint8 --global-batch 128 > gpurun_out/bench_yolo_strong_${N}gpu.log 2>&1
tail -n 1 gpurun_out/bench_yolo_strong_${N}gpu.log | cut -c1-1200
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --workload yolov5s_int8 --global-batch 64 > gpurun_out/bench_yolov5s_strong_${N}gpu.log 2>&1
tail -n 1 gpurun_out/bench_yolov5s_strong_${N}gpu.log | cut -c1-600
# the unmodified tm_benchmark on the whole group (environment selects the GPUs), batch 32 per GPU
TG_B200_GPUS=$N timeout -k 10 120 build/tengine/tm_benchmark -d B200 -m oracle/_ref/models/mobilenet_v1_int8.tmfile -i $((32*N)),3,224,224 -f 2 -r 20 -t 8 > gpurun_out/tm_benchmark_${N}gpu.log 2>&1
tail -n 3 gpurun_out/tm_benchmark_${N}gpu.log
