#!/bin/bash
# round 2, call g: halo diagnostics; with the halo path OFF: pooled-feature pyramid GEMMs + warp lookup + instnorm unroll + CUDA graph
# (tests, parity, bench), accumulation-segment experiments of the 3xF16 kernel, latency probe
mkdir -p gpurun_out
timeout 200 python scripts/halo_diag.py > gpurun_out/r02g_halo_diag.log 2>&1; grep -E "tap \(|^    " gpurun_out/r02g_halo_diag.log | head -40
echo "=== base_offset forced 0"
GIMMVFI_HALO_DBG=1 timeout 200 python scripts/halo_diag.py > gpurun_out/r02g_halo_diag_dbg1.log 2>&1; grep "tap (" gpurun_out/r02g_halo_diag_dbg1.log
export GIMMVFI_TC_HALO=0
( time timeout 1500 python -m pytest tests -m gpu -q -k "not halo" ) > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r02g_pytest_gpu.log | cut -c1-300
timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02g_parity.log 2>&1; echo "== parity"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02g_parity.log | cut -c1-420
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02g_profile.json > gpurun_out/r02g_bench.log 2>&1; tail -n 1 gpurun_out/r02g_bench.log | cut -c1-250
for sg in 2 3; do
  GIMMVFI_TC_SEG_F16=$sg timeout 600 python -m pytest tests/test_bench_parity_gpu.py -q -s > gpurun_out/r02g_parity_seg$sg.log 2>&1; echo "== parity seg_f16=$sg"; grep -E "^big_r|\.big_r|Fbig_r|passed|failed" gpurun_out/r02g_parity_seg$sg.log | cut -c1-420
  GIMMVFI_TC_SEG_F16=$sg timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02g_profile_seg$sg.json > gpurun_out/r02g_bench_seg$sg.log 2>&1; tail -n 1 gpurun_out/r02g_bench_seg$sg.log | cut -c1-250
done
timeout 300 python scripts/latency_probe.py > gpurun_out/r02g_latency.log 2>&1; cat gpurun_out/r02g_latency.log
