#!/bin/bash
# round 2, call s: what bounds a K step of the 3xF16 kernel (splitter off / drain off), and the compact epilogue against the committed one on the SAME box
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv > gpurun_out/r02s_probe.log
PROBE_GRU=1 timeout 600 python scripts/tc_split_probe.py PROBE_STALL=1 PROBE_STALL=1,PROBE_EPI=q PROBE_STALL=1,GIMMVFI_TC_DEBUG=1 PROBE_STALL=1,GIMMVFI_TC_DEBUG=2 PROBE_STALL=1,GIMMVFI_TC_DEBUG=3 >> gpurun_out/r02s_probe.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02s_profile.json > gpurun_out/r02s_bench.log 2>&1; tail -n 1 gpurun_out/r02s_bench.log | cut -c1-250
cp gimm-vfi_b200/libgimmvfi_b200.so /tmp/new.so; cp gimm-vfi_b200/libgimmvfi_b200_head.so gimm-vfi_b200/libgimmvfi_b200.so
echo "=== committed conv_tc.cu" >> gpurun_out/r02s_probe.log
PROBE_GRU=1 timeout 600 python scripts/tc_split_probe.py PROBE_STALL=1 PROBE_STALL=1,PROBE_EPI=q >> gpurun_out/r02s_probe.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02s_profile_head.json > gpurun_out/r02s_bench_head.log 2>&1; tail -n 1 gpurun_out/r02s_bench_head.log | cut -c1-250
cp /tmp/new.so gimm-vfi_b200/libgimmvfi_b200.so
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02s_bench2.log 2>&1; tail -n 1 gpurun_out/r02s_bench2.log | cut -c1-250
cut -c1-230 gpurun_out/r02s_probe.log
