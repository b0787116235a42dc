mkdir -p gpurun_out
timeout 200 python scripts/tc_microbench.py all 2>&1 | grep -v Warn | tee gpurun_out/tc_micro.log
GIMMVFI_TC_SPIN_LIMIT=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 3 -c 1 -o gpurun_out/prof_tc_split python scripts/tc_microbench.py split 1 > gpurun_out/ncu_tc_split.log 2>&1
tail -2 gpurun_out/ncu_tc_split.log
GIMMVFI_TC_SPIN_LIMIT=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 3 -c 1 -o gpurun_out/prof_tc_plain python scripts/tc_microbench.py plain 1 > gpurun_out/ncu_tc_plain.log 2>&1
tail -2 gpurun_out/ncu_tc_plain.log
