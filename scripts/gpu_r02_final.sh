#!/bin/bash
# round 2, end-of-round evidence: full GPU suite, smoke(), the default bench line (with both baselines), GIMM-VFI-F bench lines (configs 3 / 4),
# FlowFormer GRU hoisting A/B, ncu launch list of the default bench step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02fin_all.log 2>&1; echo "all rc=$?"; tail -n 3 gpurun_out/r02fin_all.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r02fin_smoke.log | cut -c1-250
timeout 900 python bench.py --profile-json gpurun_out/r02fin_r_profile.json > gpurun_out/r02fin_bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/r02fin_bench.log | cut -c1-300
timeout 600 python bench.py --config f2k --profile-json gpurun_out/r02fin_f2k_profile.json > gpurun_out/r02fin_bench_f2k.log 2>&1; echo "f2k rc=$?"; tail -n 1 gpurun_out/r02fin_bench_f2k.log | cut -c1-300
timeout 600 python bench.py --config f4k --steps 3 > gpurun_out/r02fin_bench_f4k.log 2>&1; echo "f4k rc=$?"; tail -n 1 gpurun_out/r02fin_bench_f4k.log | cut -c1-300
for i in 1 2; do
timeout 300 python scripts/f_bench.py > gpurun_out/r02fin_fbench_hoist_$i.log 2>&1; head -1 gpurun_out/r02fin_fbench_hoist_$i.log | cut -c1-200
GIMMVFI_GRU_HOIST=0 timeout 300 python scripts/f_bench.py > gpurun_out/r02fin_fbench_nohoist_$i.log 2>&1; head -1 gpurun_out/r02fin_fbench_nohoist_$i.log | cut -c1-200
done
export GIMMVFI_TC_SPIN_LIMIT=0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r02fin_launches_1080p.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline > gpurun_out/r02fin_under_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r02fin_launches_1080p.csv
