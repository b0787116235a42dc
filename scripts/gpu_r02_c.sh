#!/bin/bash
# round 2, call c: 3xF16 split kernel + fp32-class fused HypoNet: unit tests, parity at the benchmarked configs, bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -s -k "hyponet" > gpurun_out/r02c_hyponet_unit.log 2>&1; echo "hyponet unit rc=$?"; grep -E "hyponet fused|passed|failed|rror" gpurun_out/r02c_hyponet_unit.log | head -20
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -s -k "3xtf32 or 3xf16 or gru or cluster" > gpurun_out/r02c_split_unit.log 2>&1; echo "split unit rc=$?"; grep -E "split case|cluster case|passed|failed|rror" gpurun_out/r02c_split_unit.log | cut -c1-200 | head -70
for sf in 1 0; do
  GIMMVFI_TC_SPLIT_F16=$sf timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02c_parity_sf$sf.log 2>&1
  echo "== parity, 3xF16=$sf"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02c_parity_sf$sf.log | cut -c1-330
  GIMMVFI_TC_SPLIT_F16=$sf timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02c_profile_sf$sf.json > gpurun_out/r02c_bench_sf$sf.log 2>&1; tail -n 1 gpurun_out/r02c_bench_sf$sf.log | cut -c1-250
done
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r02c_pytest_gpu.log | cut -c1-300
