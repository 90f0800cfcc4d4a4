#!/bin/bash
# round 2: stage-by-stage diagnosis of the single-process multi-GPU path on a real 2-GPU box (torch-free, watchdogs, ~1 minute)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,compute_mode --format=csv,noheader > gpurun_out/multi_debug_gpus.txt 2>&1
NCCL_DEBUG=WARN TB200_DEBUG_STAGES=1 DEBUG_MULTI_WATCHDOG=20 timeout -k 5 140 python tools/debug_multi.py ${1:-2} ${2:-gpu1_alone,nccl_mobilenet,memcpy_peer_mobilenet} > gpurun_out/multi_debug.log 2>&1
tail -n 70 gpurun_out/multi_debug.log | cut -c1-420
