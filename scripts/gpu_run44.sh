#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_video_gpu.py -x -q > gpurun_out/pytest_tc.log 2>&1
tail -n 6 gpurun_out/pytest_tc.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --profile-json gpurun_out/profile_bench_default.json > gpurun_out/bench_default.log 2>&1
tail -n 2 gpurun_out/bench_default.log | cut -c1-300
timeout 600 python scripts/video_bench.py 9 > gpurun_out/video_bench.log 2>&1; cat gpurun_out/video_bench.log
