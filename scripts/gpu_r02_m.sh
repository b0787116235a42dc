#!/bin/bash
# round 2, call m: GRU context hoisting (A/B), precision mode 4 re-evaluated with the halo kernel (parity + bench)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bench_parity_gpu.py tests/test_forward_gpu.py -q -s > gpurun_out/r02m_parity.log 2>&1; echo "== parity rc=$?"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02m_parity.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02m_profile.json > gpurun_out/r02m_bench.log 2>&1; tail -n 1 gpurun_out/r02m_bench.log | cut -c1-250
GIMMVFI_GRU_HOIST=0 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02m_bench_nohoist.log 2>&1; tail -n 1 gpurun_out/r02m_bench_nohoist.log | cut -c1-250
GIMMVFI_TEST_MODE=4 timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02m_parity_mode4.log 2>&1; echo "== parity mode 4 rc=$?"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02m_parity_mode4.log | cut -c1-200
timeout 300 python bench.py --precision mixed4 --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02m_profile_mode4.json > gpurun_out/r02m_bench_mode4.log 2>&1; tail -n 1 gpurun_out/r02m_bench_mode4.log | cut -c1-250
