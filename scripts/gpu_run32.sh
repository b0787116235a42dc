#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q > gpurun_out/pytest_tc.log 2>&1
tail -n 3 gpurun_out/pytest_tc.log
timeout 900 python scripts/tc_split_probe.py > gpurun_out/split_probe.log 2>&1
cat gpurun_out/split_probe.log
