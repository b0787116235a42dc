#!/bin/bash
# round 2, call ff4: 7x7 small-cout kernel v2 (4 px / thread), volume-free RAFT lookups: parity, timings, the 4K pair
# without ds_factor that the volume path cannot hold
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tokens_ops.py tests/test_f_gpu.py -q -x -m gpu > gpurun_out/r02ff4_ops_f.log 2>&1; echo "ops+f rc=$?"; tail -n 2 gpurun_out/r02ff4_ops_f.log | cut -c1-200
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_bench_parity_gpu.py -q -x -s > gpurun_out/r02ff4_parity.log 2>&1; echo "parity rc=$?"; grep -E "volume-free|^big_r|\.big_r|passed|failed" gpurun_out/r02ff4_parity.log | cut -c1-220 | tail -12
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02ff4_r_profile.json > gpurun_out/r02ff4_bench_$i.log 2>&1; tail -n 1 gpurun_out/r02ff4_bench_$i.log | cut -c1-240
GIMMVFI_CONV7_SMALL=0 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02ff4_bench_conv7off_$i.log 2>&1; tail -n 1 gpurun_out/r02ff4_bench_conv7off_$i.log | cut -c1-240
done
GIMMVFI_RAFT_CORR_DIRECT=1 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02ff4_r_direct_profile.json > gpurun_out/r02ff4_bench_raft_direct.log 2>&1; tail -n 1 gpurun_out/r02ff4_bench_raft_direct.log | cut -c1-240
timeout 600 python bench.py --height 2176 --width 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02ff4_r_4k_profile.json > gpurun_out/r02ff4_bench_r_4k_nods.log 2>&1; echo "4k rc=$?"; tail -n 1 gpurun_out/r02ff4_bench_r_4k_nods.log | cut -c1-400
timeout 600 python scripts/f_bench.py --profile-json gpurun_out/r02ff4_f_profile.json > gpurun_out/r02ff4_fbench.log 2>&1; echo "fbench rc=$?"; head -12 gpurun_out/r02ff4_fbench.log | cut -c1-200
timeout 300 python scripts/f_bench.py --h 2176 --w 4096 --ds 0.25 --profile-json gpurun_out/r02ff4_f4k_profile.json > gpurun_out/r02ff4_fbench_4k.log 2>&1; echo "fbench4k rc=$?"; head -8 gpurun_out/r02ff4_fbench_4k.log | cut -c1-300
