mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 600 python bench.py --profile-json gpurun_out/profile_bench_default.json 2>&1 | tail -1 | tee gpurun_out/bench_default.log
