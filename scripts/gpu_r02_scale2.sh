#!/bin/bash
# round 2: two GPUs of one box under torchrun (gpurun --gpus 2): the default line (weak scaling) and GIMM-VFI-F config 3 (pairs sharded, one all-gather)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-torch-baseline 2>&1 | grep -v -i warn | tail -1 | tee gpurun_out/r02_bench_n2_final.log | cut -c1-260
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --config f2k --steps 5 --warmup 3 2>&1 | grep -v -i warn | tail -1 | tee gpurun_out/r02_bench_f2k_n2.log | cut -c1-260
