#!/bin/bash
# round 2, call y: direct-store plain epilogue instantiation of the split kernel (EPI = 2): unit, probe, parity, bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -x > gpurun_out/r02y_unit.log 2>&1; echo "unit rc=$?"; tail -n 3 gpurun_out/r02y_unit.log | cut -c1-200
PROBE_GRU=1 timeout 600 python scripts/tc_split_probe.py PROBE_STALL=1 PROBE_STALL=1,GIMMVFI_TC_DIRECT_EPI=0,GIMMVFI_TC_DIRECT_EPI_OFF=1 > gpurun_out/r02y_probe.log 2>&1; cut -c1-260 gpurun_out/r02y_probe.log
timeout 600 python -m pytest tests/test_bench_parity_gpu.py tests/test_forward_gpu.py -q -s > gpurun_out/r02y_parity.log 2>&1; echo "== parity rc=$?"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02y_parity.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02y_profile.json > gpurun_out/r02y_bench.log 2>&1; tail -n 1 gpurun_out/r02y_bench.log | cut -c1-250
GIMMVFI_TC_DIRECT_EPI=0 GIMMVFI_TC_DIRECT_EPI_OFF=1 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02y_bench_generic.log 2>&1; tail -n 1 gpurun_out/r02y_bench_generic.log | cut -c1-250
