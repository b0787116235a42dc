mkdir -p gpurun_out
timeout 200 python scripts/tc_microbench.py 2>&1 | grep -v Warn | tee gpurun_out/tc_micro.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 3 -c 1 -o gpurun_out/prof_tc python scripts/tc_microbench.py 1 > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
