mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py --profile-json gpurun_out/profile_bench_tf32.json 2>&1 | tail -1 | tee gpurun_out/bench_tf32.log
timeout 600 python bench.py --precision 3xtf32 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_3xtf32.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.log
