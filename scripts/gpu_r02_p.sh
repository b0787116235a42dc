#!/bin/bash
# round 2, call p: ncu source-level capture of the split kernel with the GRU candidate's epilogue (residual + tanh + z/h blend)
mkdir -p gpurun_out
export GIMMVFI_TC_SPIN_LIMIT=0
PROBE_GRU=1 PROBE_EPI=q timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 30 -c 1 -f -o gpurun_out/r02p_ncu_gru_q python scripts/tc_split_probe.py child > gpurun_out/r02p_ncu.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r02p_ncu.log
PROBE_GRU=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 30 -c 1 -f -o gpurun_out/r02p_ncu_gru_plain python scripts/tc_split_probe.py child > gpurun_out/r02p_ncu2.log 2>&1; echo "rc=$?"
