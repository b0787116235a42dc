#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --profile-json gpurun_out/profile_bench_default.json > gpurun_out/bench_default.log 2>&1
tail -n 2 gpurun_out/bench_default.log | cut -c1-300
