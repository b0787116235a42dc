#!/bin/bash
# round 2, last call: 16-byte form of the full-resolution up-sampling (ds_factor configurations): full GPU suite + the two GIMM-VFI-F lines + default line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02fin4_all.log 2>&1; echo "all rc=$?"; tail -n 3 gpurun_out/r02fin4_all.log | cut -c1-200
timeout 300 python bench.py --config f4k --steps 3 --profile-json gpurun_out/r02fin4_f4k_profile.json > gpurun_out/r02fin4_bench_f4k.log 2>&1; tail -n 1 gpurun_out/r02fin4_bench_f4k.log | cut -c1-260
timeout 300 python bench.py --config f2k --profile-json gpurun_out/r02fin4_f2k_profile.json > gpurun_out/r02fin4_bench_f2k.log 2>&1; tail -n 1 gpurun_out/r02fin4_bench_f2k.log | cut -c1-260
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02fin4_bench.log 2>&1; tail -n 1 gpurun_out/r02fin4_bench.log | cut -c1-260
