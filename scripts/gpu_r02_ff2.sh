#!/bin/bash
# round 2, FlowFormer call 2: shared-memory K/V attention + constant-bank cost_conv1: parity, config-3 timing / profile, ncu launch list and
# --set full captures of the new HBM / CUDA-core kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_f_gpu.py -q -x -s > gpurun_out/r02ff2_f.log 2>&1; echo "f rc=$?"; grep -E "mode|imgt_pred|passed|failed|Error|error" gpurun_out/r02ff2_f.log | cut -c1-220 | tail -30
timeout 600 python scripts/f_bench.py --profile-json gpurun_out/r02ff2_f_profile.json > gpurun_out/r02ff2_fbench.log 2>&1; echo "fbench rc=$?"; head -32 gpurun_out/r02ff2_fbench.log | cut -c1-200
timeout 300 python scripts/f_bench.py --h 2176 --w 4096 --ds 0.25 > gpurun_out/r02ff2_fbench_4k.log 2>&1; echo "fbench4k rc=$?"; head -3 gpurun_out/r02ff2_fbench_4k.log | cut -c1-300
export GIMMVFI_TC_SPIN_LIMIT=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02ff2_launches_f.csv \
   python scripts/f_bench.py --steps 1 --warmup 1 > gpurun_out/r02ff2_under_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r02ff2_launches_f.csv
for spec in "attention_shared_kv:0:2:attn" "CostConv1K:0:1:costconv1" "WindowAttnK:0:1:winattn" "layernorm_warp:30:1:layernorm" "StridedAttnKILi16:0:1:attn16" "PatchifyK4:3:1:patchify"; do
  IFS=: read -r pat skip cnt tag <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s "$skip" -c "$cnt" -f -o gpurun_out/r02ff2_ncu_$tag \
     python scripts/f_bench.py --steps 1 --warmup 0 > gpurun_out/r02ff2_under_ncu_$tag.log 2>&1; echo "ncu $tag rc=$?"
done
ls -la gpurun_out/*.ncu-rep
