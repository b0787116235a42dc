#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench step, (2) --set full of the two dominant tensor-core kernels
mkdir -p gpurun_out
export GIMMVFI_TC_SPIN_LIMIT=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_mode3.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_mode3.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 6 -c 2 -f -o gpurun_out/prof_tc_f16 \
   python scripts/tc_f16_probe.py > gpurun_out/ncu_f16.log 2>&1
echo "f16 rc=$?"
PROBE_ONLY_GRU=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 5 -c 1 -f -o gpurun_out/prof_tc_split \
   python scripts/tc_microbench.py split 1 > gpurun_out/ncu_split.log 2>&1
echo "split rc=$?"
ls -la gpurun_out/*.ncu-rep
