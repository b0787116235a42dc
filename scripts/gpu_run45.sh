#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_video_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/pytest_tc.log 2>&1
tail -n 4 gpurun_out/pytest_tc.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --profile-json gpurun_out/profile_bench_default.json > gpurun_out/bench_default.log 2>&1
tail -n 2 gpurun_out/bench_default.log | cut -c1-300
GIMMVFI_TC_SPIN_LIMIT=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1
echo "memcheck rc=$?"; tail -n 6 gpurun_out/memcheck_smoke.log
