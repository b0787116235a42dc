mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -rA -s 2>&1 | grep -E "^case|PASSED|FAILED|passed|failed|Error|error" | head -60 > gpurun_out/pytest_tc.log; cat gpurun_out/pytest_tc.log
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -k softsplat 2>&1 | tail -3
timeout 400 python scripts/tc_e2e_check.py 2>&1 | grep -v Warn | tail -40 | tee gpurun_out/tc_e2e.log
timeout 600 python bench.py --steps 3 --warmup 3 --profile-json gpurun_out/profile_1080p.json 2>&1 | tail -2 | tee gpurun_out/bench.log
