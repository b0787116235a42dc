#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/tc_split_probe.py > gpurun_out/split_probe.log 2>&1
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q > gpurun_out/pytest_tc.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --profile-json gpurun_out/profile_bench_default.json > gpurun_out/bench_default.log 2>&1
tail -3 gpurun_out/pytest_tc.log gpurun_out/pytest_gpu.log; cat gpurun_out/split_probe.log; tail -2 gpurun_out/bench_default.log
