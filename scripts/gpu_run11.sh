mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -rA -s 2>&1 | grep -E "split case|FAILED|passed|failed|rror:" | head -40 | tee gpurun_out/pytest_tc.log
timeout 600 python -m pytest tests/test_forward_gpu.py -q -rA -s 2>&1 | grep -E " mode |FAILED|passed|failed|rror" | cut -c1-900 | tee gpurun_out/pytest_fwd.log
timeout 900 python scripts/parity_1080p.py 2>&1 | grep -v Warn | tee gpurun_out/parity_1080p.log
timeout 500 python scripts/tc_e2e_check.py 2>&1 | grep -v Warn | grep -A22 "1080p tc=2" | tee gpurun_out/tc_e2e.log
