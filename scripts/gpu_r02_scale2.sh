#!/bin/bash
# 2 GPUs (gpurun --gpus 2): the headline workload weak-scaled, and BASELINE config 5 (256 x 1280x720 pairs) strong-scaled, under torchrun
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02w_bench_pair1080_n2.log 2>&1; tail -n 1 gpurun_out/r02w_bench_pair1080_n2.log | cut -c1-300
timeout 900 python bench.py --config batch720 --pairs 256 --micro-batch 8 --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline > gpurun_out/r02w_bench_batch720_n1.log 2>&1; tail -n 1 gpurun_out/r02w_bench_batch720_n1.log | cut -c1-300
timeout 900 $TR --master-port 29542 bench.py --gpus 2 --config batch720 --pairs 256 --micro-batch 8 --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline > gpurun_out/r02w_bench_batch720_n2.log 2>&1; tail -n 1 gpurun_out/r02w_bench_batch720_n2.log | cut -c1-300
