#!/bin/bash
# round 2, call u: is the 3xF16 K step bound by the tensor pipe?  (one MMA term instead of three; with and without splitter / drain work)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -x > gpurun_out/r02u_unit.log 2>&1; echo "unit rc=$?"; tail -n 3 gpurun_out/r02u_unit.log | cut -c1-200
GIMMVFI_TC_SPLIT_HALO=1 timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -x -k "cluster or gru" > gpurun_out/r02u_unit_halo.log 2>&1; echo "unit halo rc=$?"; tail -n 3 gpurun_out/r02u_unit_halo.log | cut -c1-200
PROBE_GRU=1 timeout 600 python scripts/tc_split_probe.py PROBE_STALL=1 PROBE_STALL=1,GIMMVFI_TC_DEBUG=4 PROBE_STALL=1,GIMMVFI_TC_DEBUG=7 PROBE_STALL=1,GIMMVFI_TC_DEBUG=7,GIMMVFI_TC_SPLIT_HALO=1 > gpurun_out/r02u_probe.log 2>&1; cut -c1-260 gpurun_out/r02u_probe.log
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02u_profile.json > gpurun_out/r02u_bench.log 2>&1; tail -n 1 gpurun_out/r02u_bench.log | cut -c1-250
GIMMVFI_TC_SPLIT_HALO=1 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02u_bench_halo.log 2>&1; tail -n 1 gpurun_out/r02u_bench_halo.log | cut -c1-250
timeout 300 python bench.py --precision mixed4 --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02u_profile_mode4.json > gpurun_out/r02u_bench_mode4.log 2>&1; tail -n 1 gpurun_out/r02u_bench_mode4.log | cut -c1-250
