#!/bin/bash
# first GPU run for the experimental precision mode 4 (fp16 storage of the K-poor chains): end-to-end parity at the golden sizes,
# 1080p parity vs the CPU oracle, and the bench line next to the default mode's
mkdir -p gpurun_out
GIMMVFI_TEST_MODE4=1 timeout 900 python -m pytest tests/test_forward_gpu.py -x -q -k "f16_chains" > gpurun_out/pytest_mode4.log 2>&1; tail -n 5 gpurun_out/pytest_mode4.log
timeout 300 python bench.py --no-cpu-baseline --precision mixed4 --profile-json gpurun_out/profile_bench_mode4.json > gpurun_out/bench_mode4.log 2>&1; tail -n 1 gpurun_out/bench_mode4.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_default.log 2>&1; tail -n 1 gpurun_out/bench_default.log | cut -c1-300
PARITY_MODES=3,4 timeout 900 python scripts/parity_1080p.py > gpurun_out/parity_1080p_mode4.log 2>&1; cat gpurun_out/parity_1080p_mode4.log
