#!/bin/bash
# round 2, call v: swizzled epilogue staging (4 pipeline stages in the 3xF16 kernel at N = 128); mode 4 as the default; full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -x > gpurun_out/r02v_unit.log 2>&1; echo "unit rc=$?"; tail -n 3 gpurun_out/r02v_unit.log | cut -c1-200
PROBE_GRU=1 timeout 600 python scripts/tc_split_probe.py PROBE_STALL=1 PROBE_STALL=1,PROBE_EPI=q > gpurun_out/r02v_probe.log 2>&1; cut -c1-260 gpurun_out/r02v_probe.log
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02v_profile.json > gpurun_out/r02v_bench.log 2>&1; tail -n 1 gpurun_out/r02v_bench.log | cut -c1-250
timeout 300 python bench.py --precision mixed3 --no-cpu-baseline --no-torch-baseline > gpurun_out/r02v_bench_mode3.log 2>&1; tail -n 1 gpurun_out/r02v_bench_mode3.log | cut -c1-250
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02v_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -n 5 gpurun_out/r02v_pytest_gpu.log | cut -c1-250
