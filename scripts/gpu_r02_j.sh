#!/bin/bash
# round 2, call j: halo kernel with the direct-store epilogue, x-packed 5x5, bench; compute-sanitizer racecheck / synccheck / memcheck of smoke()
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py tests/test_kernels_gpu.py -q -k "halo or softsplat" > gpurun_out/r02j_unit.log 2>&1; echo "unit rc=$?"; tail -n 3 gpurun_out/r02j_unit.log | cut -c1-200
timeout 300 python scripts/halo_probe.py > gpurun_out/r02j_halo_probe.log 2>&1; cat gpurun_out/r02j_halo_probe.log
timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02j_parity.log 2>&1; echo "== parity"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02j_parity.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02j_profile.json > gpurun_out/r02j_bench.log 2>&1; tail -n 1 gpurun_out/r02j_bench.log | cut -c1-250
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 4 gpurun_out/r02j_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -n 4 gpurun_out/r02j_sanitizer_synccheck.log
timeout 1200 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 6 gpurun_out/r02j_sanitizer_racecheck.log | cut -c1-300
