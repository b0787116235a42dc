#!/bin/bash
# round 2, first GPU call: full GPU test suite (incl. reference-generated fixtures at the benchmarked sizes), default bench with the
# torch-GPU baseline, mode-4 validation on hardware, HBM-kernel probe + ncu --set full of those kernels
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r02a_pytest_gpu.log
timeout 600 python bench.py --profile-json gpurun_out/r02a_profile_default.json > gpurun_out/r02a_bench_default.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/r02a_bench_default.log | cut -c1-400
GIMMVFI_TEST_MODE=4 timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02a_parity_mode4.log 2>&1; echo "mode4 parity rc=$?"; grep -E "big_r|passed|failed" gpurun_out/r02a_parity_mode4.log | cut -c1-600
timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02a_parity_mode3.log 2>&1; grep -E "big_r|passed|failed" gpurun_out/r02a_parity_mode3.log | cut -c1-600
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --precision mixed4 --profile-json gpurun_out/r02a_profile_mode4.json > gpurun_out/r02a_bench_mode4.log 2>&1; tail -n 1 gpurun_out/r02a_bench_mode4.log | cut -c1-300
timeout 300 python scripts/hbm_kernels_probe.py > gpurun_out/r02a_hbm_probe.log 2>&1; cat gpurun_out/r02a_hbm_probe.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k regex:'softsplat_acc|SplatNorm|Backwarp|ResizeK|InApply|InPartial|ConvexUp|CorrLookup' -c 12 -f -o gpurun_out/r02a_hbm_kernels \
   python scripts/hbm_kernels_probe.py --once > gpurun_out/r02a_ncu_hbm.log 2>&1; echo "ncu rc=$?"; tail -n 3 gpurun_out/r02a_ncu_hbm.log
ls -la gpurun_out/*.ncu-rep
