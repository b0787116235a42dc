mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -rA -s 2>&1 | grep -E "^case|FAILED|passed|failed|Error" | head -40 > gpurun_out/pytest_tc.log; cat gpurun_out/pytest_tc.log
timeout 600 python -m pytest tests -m gpu -q -rA -s 2>&1 | grep -E "imgt_pred max|flowt max|PASSED|FAILED|passed|failed|Error|assert" | head -80 > gpurun_out/pytest_gpu.log; tail -45 gpurun_out/pytest_gpu.log
timeout 400 python scripts/tc_e2e_check.py 2>&1 | grep -v Warn | tail -60 | tee gpurun_out/tc_e2e.log
