#!/bin/bash
# round 2, call l: volume-free BidirCorrBlock lookup (unit, parity at the benchmarked configs, bench A/B), then the ncu evidence of call k
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "corr" > gpurun_out/r02l_unit.log 2>&1; echo "unit rc=$?"; tail -n 3 gpurun_out/r02l_unit.log | cut -c1-200
timeout 600 python -m pytest tests/test_bench_parity_gpu.py tests/test_forward_gpu.py -q -s > gpurun_out/r02l_parity.log 2>&1; echo "== parity rc=$?"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02l_parity.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02l_profile.json > gpurun_out/r02l_bench.log 2>&1; tail -n 1 gpurun_out/r02l_bench.log | cut -c1-250
GIMMVFI_CORR_DIRECT=0 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02l_bench_volume.log 2>&1; tail -n 1 gpurun_out/r02l_bench_volume.log | cut -c1-250
bash scripts/gpu_r02_k.sh
