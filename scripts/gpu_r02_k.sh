#!/bin/bash
# round 2, call k: ncu evidence for profiles/: launch list of one bench step; --set full of the f16 trunk conv, the halo kernel, the fused
# HypoNet kernels, the lookup kernel (all through the probes)
mkdir -p gpurun_out
export GIMMVFI_TC_SPIN_LIMIT=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r02k_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline > gpurun_out/r02k_bench_under_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02k_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 6 -c 1 -f -o gpurun_out/r02k_ncu_f16_trunk python scripts/tc_f16_probe.py > gpurun_out/r02k_ncu_a.log 2>&1; echo "f16 trunk rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -c 1 -f -o gpurun_out/r02k_ncu_halo python scripts/halo_probe.py --once > gpurun_out/r02k_ncu_b.log 2>&1; echo "halo rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hyponet_fused -c 2 -f -o gpurun_out/r02k_ncu_hyponet python scripts/hyponet_probe.py --once > gpurun_out/r02k_ncu_c.log 2>&1; echo "hyponet rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'corr_lookup_warp|corr_lookup_direct|softsplat_acc' -c 3 -f -o gpurun_out/r02k_ncu_lookup python scripts/hbm_kernels_probe.py --once > gpurun_out/r02k_ncu_d.log 2>&1; echo "lookup rc=$?"
ls -la gpurun_out/r02k*
