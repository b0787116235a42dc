#!/bin/bash
# round 2, call z: compact plain-epilogue instantiation of the split kernel (EPI = 2, staged stores): unit, parity, bench A/B (twice, interleaved)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -x > gpurun_out/r02z_unit.log 2>&1; echo "unit rc=$?"; tail -n 3 gpurun_out/r02z_unit.log | cut -c1-200
timeout 600 python -m pytest tests/test_bench_parity_gpu.py tests/test_forward_gpu.py -q -s > gpurun_out/r02z_parity.log 2>&1; echo "== parity rc=$?"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02z_parity.log | cut -c1-200
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02z_profile.json > gpurun_out/r02z_bench_$i.log 2>&1; tail -n 1 gpurun_out/r02z_bench_$i.log | cut -c1-250
GIMMVFI_TC_PLAIN_EPI=0 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02z_bench_generic_$i.log 2>&1; tail -n 1 gpurun_out/r02z_bench_generic_$i.log | cut -c1-250
done
