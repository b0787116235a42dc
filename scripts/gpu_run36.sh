#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q > gpurun_out/pytest_tc.log 2>&1
tail -n 3 gpurun_out/pytest_tc.log
timeout 300 python scripts/tc_f16_probe.py > gpurun_out/f16_probe.log 2>&1; cat gpurun_out/f16_probe.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --profile-json gpurun_out/profile_bench_default.json > gpurun_out/bench_default.log 2>&1
tail -n 2 gpurun_out/bench_default.log | cut -c1-300
