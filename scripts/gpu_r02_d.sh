#!/bin/bash
# round 2, call d: 3xF16 correlation GEMM, RAFT-encoder TF32 experiment (stage mask 32), micro-benchmarks, ncu --set full of the tensor-core kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_kernels_gpu.py -q -s -k "corr" > gpurun_out/r02d_corr_unit.log 2>&1; echo "corr unit rc=$?"; tail -n 3 gpurun_out/r02d_corr_unit.log
timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02d_parity.log 2>&1; echo "== parity"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02d_parity.log | cut -c1-400
GIMMVFI_PRECISE=32 timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02d_parity_k32.log 2>&1; echo "== parity, RAFT encoders TF32"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02d_parity_k32.log | cut -c1-400
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02d_profile.json > gpurun_out/r02d_bench.log 2>&1; tail -n 1 gpurun_out/r02d_bench.log | cut -c1-250
GIMMVFI_PRECISE=32 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02d_profile_k32.json > gpurun_out/r02d_bench_k32.log 2>&1; tail -n 1 gpurun_out/r02d_bench_k32.log | cut -c1-250
timeout 300 python scripts/tc_microbench.py all > gpurun_out/r02d_tc_microbench.log 2>&1; cat gpurun_out/r02d_tc_microbench.log
timeout 300 python scripts/hyponet_probe.py > gpurun_out/r02d_hyponet_probe.log 2>&1; cat gpurun_out/r02d_hyponet_probe.log
export GIMMVFI_TC_SPIN_LIMIT=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02d_ncu_3xf16_gru python scripts/tc_microbench.py split16 1 > gpurun_out/r02d_ncu_a.log 2>&1; echo "ncu 3xf16 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hyponet_fused -c 2 -f -o gpurun_out/r02d_ncu_hyponet python scripts/hyponet_probe.py --once > gpurun_out/r02d_ncu_b.log 2>&1; echo "ncu hyponet rc=$?"
ls -la gpurun_out/*.ncu-rep
