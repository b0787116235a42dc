mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -v -i warn | tail -2 | tee gpurun_out/bench_n2.log
