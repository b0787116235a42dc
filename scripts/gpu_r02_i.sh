#!/bin/bash
# round 2, call i: one-pass splat, near-final default bench (all legs), halo kernel ncu, config-5 batch on one GPU, the same-config reference arm once
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_conv_tc_gpu.py tests/test_forward_gpu.py -q -k "softsplat or halo or cuda_graph or forward_matches_reference" > gpurun_out/r02i_unit.log 2>&1; echo "unit rc=$?"; tail -n 4 gpurun_out/r02i_unit.log | cut -c1-300
timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02i_parity.log 2>&1; echo "== parity"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02i_parity.log | cut -c1-330
timeout 300 python scripts/hbm_kernels_probe.py > gpurun_out/r02i_hbm_probe.log 2>&1; cat gpurun_out/r02i_hbm_probe.log
timeout 900 python bench.py --profile-json gpurun_out/r02i_profile.json > gpurun_out/r02i_bench.log 2>&1; tail -n 1 gpurun_out/r02i_bench.log | cut -c1-600
timeout 300 python scripts/halo_probe.py > gpurun_out/r02i_halo_probe.log 2>&1; cat gpurun_out/r02i_halo_probe.log
export GIMMVFI_TC_SPIN_LIMIT=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -c 1 -f -o gpurun_out/r02i_ncu_halo python scripts/halo_probe.py --once > gpurun_out/r02i_ncu_halo.log 2>&1; echo "ncu halo rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:softsplat_tile -c 1 -f -o gpurun_out/r02i_ncu_splat python scripts/hbm_kernels_probe.py --once > gpurun_out/r02i_ncu_splat.log 2>&1; echo "ncu splat rc=$?"
unset GIMMVFI_TC_SPIN_LIMIT
timeout 600 python bench.py --config batch720 --pairs 64 --micro-batch 8 --steps 2 --warmup 1 > gpurun_out/r02i_bench_batch720_n1.log 2>&1; tail -n 1 gpurun_out/r02i_bench_batch720_n1.log | cut -c1-500
timeout 600 python bench.py --config batch720 --pairs 64 --micro-batch 1 --steps 2 --warmup 1 > gpurun_out/r02i_bench_batch720_n1_mb1.log 2>&1; tail -n 1 gpurun_out/r02i_bench_batch720_n1_mb1.log | cut -c1-300
GIMMVFI_REF_MAX_STEPS=1 GIMMVFI_REF_MAX_WARMUP=0 timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02i_bench_reference_arm.log 2>&1; tail -n 1 gpurun_out/r02i_bench_reference_arm.log | cut -c1-700
