#!/bin/bash
# short confirmation run: the whole GPU test suite + the default bench line without the CPU baseline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_default.log 2>&1
tail -n 1 gpurun_out/bench_default.log | cut -c1-260
