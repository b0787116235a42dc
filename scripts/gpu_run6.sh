mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -rA 2>&1 | grep -E "FAILED|passed|failed|Error" | head -20 | tee gpurun_out/pytest_tc.log
timeout 200 python scripts/tc_microbench.py 2>&1 | grep -v Warn | tee gpurun_out/tc_micro.log
timeout 400 python scripts/tc_e2e_check.py 2>&1 | grep -v Warn | tail -48 | tee gpurun_out/tc_e2e.log
