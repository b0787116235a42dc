#!/bin/bash
# round 2, call b: fused HypoNet unit test; which post-RAFT stage limits parity on the reference's demo frames (stage mask sweep)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -s -k "hyponet or softsplat" > gpurun_out/r02b_hyponet_unit.log 2>&1; echo "hyponet unit rc=$?"; grep -E "hyponet fused|passed|failed|Error|error" gpurun_out/r02b_hyponet_unit.log | head -20
for mode in 0 1 3; do
  GIMMVFI_TEST_MODE=$mode timeout 300 python -m pytest tests/test_bench_parity_gpu.py -q -s -k "demo or 736x1280" > gpurun_out/r02b_demo_mode$mode.log 2>&1
  echo "== mode $mode"; grep -E "^big_r|\.big_r|Fbig_r" gpurun_out/r02b_demo_mode$mode.log | cut -c1-330
done
for kn in 1 2 3 4 7 8 15 31; do
  GIMMVFI_PRECISE=$kn timeout 300 python -m pytest tests/test_bench_parity_gpu.py -q -s -k "demo or 736x1280" > gpurun_out/r02b_demo_knob$kn.log 2>&1
  echo "== mode 3 + stage mask $kn"; grep -E "^big_r|\.big_r|Fbig_r" gpurun_out/r02b_demo_knob$kn.log | cut -c1-330
done
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02b_profile.json > gpurun_out/r02b_bench.log 2>&1; tail -n 1 gpurun_out/r02b_bench.log | cut -c1-300
