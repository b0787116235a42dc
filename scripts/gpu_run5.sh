mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -rA -s 2>&1 | grep -E "^case|FAILED|passed|failed|Error" | head -40 > gpurun_out/pytest_tc.log; cat gpurun_out/pytest_tc.log
timeout 200 python scripts/tc_microbench.py 2>&1 | grep -v Warn | tee gpurun_out/tc_micro.log
timeout 400 python scripts/tc_e2e_check.py 2>&1 | grep -v Warn | tail -48 | tee gpurun_out/tc_e2e.log
timeout 300 ncu --set full --replay-mode application --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 3 -c 1 -o gpurun_out/prof_tc python scripts/tc_microbench.py 1 > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
