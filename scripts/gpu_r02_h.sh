#!/bin/bash
# round 2, call h: halo kernel (fixed descriptors; pitch 16 / 10 / three copies), full GPU suite, full-frame parity (accumulation segments), latency
mkdir -p gpurun_out
for hm in 1 2 3; do
  GIMMVFI_HALO_MODE=$hm timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -s -k "halo" > gpurun_out/r02h_halo_unit_m$hm.log 2>&1; echo "halo mode $hm unit rc=$?"; grep -E "halo case|passed|failed" gpurun_out/r02h_halo_unit_m$hm.log | cut -c1-160
done
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r02h_pytest_gpu.log | cut -c1-300
for hm in 1 2 0; do
  export GIMMVFI_HALO_MODE=$hm; export GIMMVFI_TC_HALO=$([ $hm = 0 ] && echo 0 || echo 1)
  timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02h_parity_halo$hm.log 2>&1; echo "== parity halo mode $hm"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02h_parity_halo$hm.log | cut -c1-200
  timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02h_profile_halo$hm.json > gpurun_out/r02h_bench_halo$hm.log 2>&1; tail -n 1 gpurun_out/r02h_bench_halo$hm.log | cut -c1-250
done
unset GIMMVFI_HALO_MODE GIMMVFI_TC_HALO
timeout 400 python scripts/parity_fullframe.py --make-ref > gpurun_out/r02h_fullframe.log 2>&1
for sg in 1 2; do GIMMVFI_TC_SEG_F16=$sg timeout 200 python scripts/parity_fullframe.py >> gpurun_out/r02h_fullframe.log 2>&1; done
GIMMVFI_HYPO_FAST=1 timeout 200 python scripts/parity_fullframe.py >> gpurun_out/r02h_fullframe.log 2>&1
cat gpurun_out/r02h_fullframe.log | cut -c1-330
timeout 300 python scripts/latency_probe.py > gpurun_out/r02h_latency.log 2>&1; cat gpurun_out/r02h_latency.log
