#!/bin/bash
# end-of-round evidence run (1 GPU): tests, smoke, the default bench line (with the bounded CPU baseline), 1080p parity, clip bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --profile-json gpurun_out/profile_bench_default.json > gpurun_out/bench_default_full.log 2>&1
tail -n 1 gpurun_out/bench_default_full.log | cut -c1-400
PARITY_MODES=3 timeout 900 python scripts/parity_1080p.py > gpurun_out/parity_1080p.log 2>&1
cat gpurun_out/parity_1080p.log
timeout 600 python scripts/video_bench.py 9 > gpurun_out/video_bench.log 2>&1; tail -n 2 gpurun_out/video_bench.log
