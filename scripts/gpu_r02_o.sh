#!/bin/bash
# round 2, call o: side-tensor epilogue (L1 prefetch + batched loads) A/B against the sequential order; mode 4 with the zero-K MMAs skipped
mkdir -p gpurun_out
L=gimm-vfi_b200/libgimmvfi_b200.so
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -x > gpurun_out/r02o_unit.log 2>&1; echo "unit rc=$?"; tail -n 2 gpurun_out/r02o_unit.log | cut -c1-200
PROBE_GRU=1 timeout 600 python scripts/tc_split_probe.py PROBE_STALL=1 PROBE_STALL=1,PROBE_EPI=q > gpurun_out/r02o_gru_probe_batched.log 2>&1; cut -c1-260 gpurun_out/r02o_gru_probe_batched.log
timeout 600 python -m pytest tests/test_bench_parity_gpu.py tests/test_forward_gpu.py -q -s > gpurun_out/r02o_parity.log 2>&1; echo "== parity rc=$?"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02o_parity.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02o_profile.json > gpurun_out/r02o_bench.log 2>&1; tail -n 1 gpurun_out/r02o_bench.log | cut -c1-250
GIMMVFI_GRU_HOIST=0 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02o_bench_nohoist.log 2>&1; tail -n 1 gpurun_out/r02o_bench_nohoist.log | cut -c1-250
GIMMVFI_TEST_MODE=4 timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02o_parity_mode4.log 2>&1; echo "== parity mode 4 rc=$?"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02o_parity_mode4.log | cut -c1-200
timeout 300 python bench.py --precision mixed4 --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02o_profile_mode4.json > gpurun_out/r02o_bench_mode4.log 2>&1; tail -n 1 gpurun_out/r02o_bench_mode4.log | cut -c1-250
# --- the sequential-order build of the epilogue
cp gimm-vfi_b200/libgimmvfi_b200_seq.so $L
PROBE_GRU=1 timeout 600 python scripts/tc_split_probe.py PROBE_STALL=1,PROBE_EPI=q > gpurun_out/r02o_gru_probe_seq.log 2>&1; cut -c1-260 gpurun_out/r02o_gru_probe_seq.log
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02o_bench_seq.log 2>&1; tail -n 1 gpurun_out/r02o_bench_seq.log | cut -c1-250
