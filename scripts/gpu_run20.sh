mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --profile-json gpurun_out/profile_bench_default.json 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/bench_default.log
