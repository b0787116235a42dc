#!/bin/bash
# round 2, call q: out-of-line side-tensor epilogue (straight-line path), halo kernel with two interleaved MMA chains
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -x > gpurun_out/r02q_unit.log 2>&1; echo "unit rc=$?"; tail -n 2 gpurun_out/r02q_unit.log | cut -c1-200
timeout 300 python scripts/halo_probe.py > gpurun_out/r02q_halo_probe.log 2>&1; cat gpurun_out/r02q_halo_probe.log
PROBE_GRU=1 timeout 600 python scripts/tc_split_probe.py PROBE_STALL=1 PROBE_STALL=1,PROBE_EPI=q > gpurun_out/r02q_gru_probe.log 2>&1; cut -c1-260 gpurun_out/r02q_gru_probe.log
timeout 600 python -m pytest tests/test_bench_parity_gpu.py tests/test_forward_gpu.py -q -s > gpurun_out/r02q_parity.log 2>&1; echo "== parity rc=$?"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02q_parity.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02q_profile.json > gpurun_out/r02q_bench.log 2>&1; tail -n 1 gpurun_out/r02q_bench.log | cut -c1-250
GIMMVFI_GRU_HOIST=0 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02q_bench_nohoist.log 2>&1; tail -n 1 gpurun_out/r02q_bench_nohoist.log | cut -c1-250
timeout 300 python bench.py --precision mixed4 --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02q_profile_mode4.json > gpurun_out/r02q_bench_mode4.log 2>&1; tail -n 1 gpurun_out/r02q_bench_mode4.log | cut -c1-250
