#!/bin/bash
# round 2, call aa: corr GEMM on the plain-epilogue instantiation: unit (corr tests), bench A/B interleaved
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_kernels_gpu.py -q -x -k "corr" > gpurun_out/r02aa_unit.log 2>&1; echo "unit rc=$?"; tail -n 3 gpurun_out/r02aa_unit.log | cut -c1-200
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02aa_profile.json > gpurun_out/r02aa_bench_$i.log 2>&1; tail -n 1 gpurun_out/r02aa_bench_$i.log | cut -c1-200
GIMMVFI_TC_PLAIN_EPI=0 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02aa_profile_generic.json > gpurun_out/r02aa_bench_generic_$i.log 2>&1; tail -n 1 gpurun_out/r02aa_bench_generic_$i.log | cut -c1-200
done
timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02aa_parity.log 2>&1; echo "== parity rc=$?"; grep -E "passed|failed" gpurun_out/r02aa_parity.log | cut -c1-200
