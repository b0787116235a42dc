#!/bin/bash
# round 2, call e: halo-reuse 3x3 conv kernel (unit + A/B bench), HypoNet with shared-space loads, GIMM-VFI-F synthesis half, 3xF16 stall probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -s -k "halo" > gpurun_out/r02e_halo_unit.log 2>&1; echo "halo unit rc=$?"; grep -E "halo case|passed|failed|rror" gpurun_out/r02e_halo_unit.log | cut -c1-200 | head -20
timeout 600 python -m pytest tests/test_f_gpu.py tests/test_kernels_gpu.py -q -s -k "f_ or hyponet" > gpurun_out/r02e_f_unit.log 2>&1; echo "F + hyponet rc=$?"; grep -E "^f_|\.f_|hyponet fused|passed|failed|rror" gpurun_out/r02e_f_unit.log | cut -c1-200 | head -30
timeout 300 python scripts/hyponet_probe.py > gpurun_out/r02e_hyponet_probe.log 2>&1; cat gpurun_out/r02e_hyponet_probe.log
for hl in 1 0; do
  GIMMVFI_TC_HALO=$hl timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02e_parity_halo$hl.log 2>&1
  echo "== parity, halo=$hl"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02e_parity_halo$hl.log | cut -c1-200
  GIMMVFI_TC_HALO=$hl timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02e_profile_halo$hl.json > gpurun_out/r02e_bench_halo$hl.log 2>&1; tail -n 1 gpurun_out/r02e_bench_halo$hl.log | cut -c1-250
done
timeout 300 python scripts/tc_split_probe.py > gpurun_out/r02e_split_probe.log 2>&1; cat gpurun_out/r02e_split_probe.log | cut -c1-1500
