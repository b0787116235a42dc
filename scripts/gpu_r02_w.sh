#!/bin/bash
# round 2, call w: end-of-round evidence on the final code: full bench line (with the CPU and stock-torch legs), ncu launch list of bench steps, compute-sanitizer on smoke()
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -x -k "halo" > gpurun_out/r02w_unit_halo.log 2>&1; echo "halo unit rc=$?"; tail -n 2 gpurun_out/r02w_unit_halo.log | cut -c1-200
timeout 900 python bench.py --profile-json gpurun_out/r02w_profile.json > gpurun_out/r02w_bench_full.log 2>&1; tail -n 1 gpurun_out/r02w_bench_full.log | cut -c1-400
GIMMVFI_TC_SPIN_LIMIT=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/r02w_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline > gpurun_out/r02w_bench_under_ncu.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r02w_launches.csv
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 3 gpurun_out/r02w_sanitizer_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -n 3 gpurun_out/r02w_sanitizer_synccheck.log | cut -c1-200
timeout 1200 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 4 gpurun_out/r02w_sanitizer_racecheck.log | cut -c1-300
timeout 600 python scripts/hbm_kernels_probe.py > gpurun_out/r02w_hbm_kernels_probe.log 2>&1; cat gpurun_out/r02w_hbm_kernels_probe.log
