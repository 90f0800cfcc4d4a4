#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_pack3.log 2>&1
TB200_DW_NO_PACK3=1 timeout 600 python bench.py --steps 20 --warmup 3 --cpu-images 0 > gpurun_out/bench_nopack3.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --ref-images 64 > gpurun_out/bench_ref_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_2gpu.log
tail -n 3 gpurun_out/pytest.log; for f in bench_pack3 bench_nopack3 bench_2gpu bench_ref_2gpu; do echo == $f; tail -n 2 gpurun_out/$f.log | cut -c1-700; done
