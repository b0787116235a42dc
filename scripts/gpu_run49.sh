#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --timesteps 7 --steps 3 --warmup 2 > gpurun_out/bench_T7.log 2>&1
tail -n 1 gpurun_out/bench_T7.log | cut -c1-400
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "thin or narrow" > gpurun_out/pytest_thin.log 2>&1; tail -n 2 gpurun_out/pytest_thin.log
