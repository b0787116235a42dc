mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -rA -s 2>&1 | grep -E "^split case|FAILED|passed|failed|Error|rror:" | head -40 | tee gpurun_out/pytest_tc.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -q 2>&1 | tail -2
timeout 500 python scripts/tc_e2e_check.py 2>&1 | grep -v Warn | grep -v "conv2d_simt_n64 k\|   conv2d_tc_tf32" | tail -50 | tee gpurun_out/tc_e2e.log
