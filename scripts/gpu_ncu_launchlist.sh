#!/bin/bash
# ncu launch list of one bench step (per-launch gpu__time_duration), final round-1 code
mkdir -p gpurun_out
export GIMMVFI_TC_SPIN_LIMIT=0
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_mode3.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_mode3.csv
