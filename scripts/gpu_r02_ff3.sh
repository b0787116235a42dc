#!/bin/bash
# round 2, FlowFormer call 3: vectorised LayerNorm + shared-memory 7x7 small-cout conv; per-op tests, F parity, full suite, R bench, F timing at
# configs 3 / 4, ncu --set full of the CUDA-core kernels, compute-sanitizer on a small F forward
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tokens_ops.py tests/test_f_gpu.py -q -x -s -m gpu > gpurun_out/r02ff3_f.log 2>&1; echo "ops+f rc=$?"; grep -E "mode|imgt_pred|passed|failed|Error|error" gpurun_out/r02ff3_f.log | cut -c1-220 | tail -24
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02ff3_all.log 2>&1; echo "all rc=$?"; tail -n 4 gpurun_out/r02ff3_all.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --profile-json gpurun_out/r02ff3_r_profile.json > gpurun_out/r02ff3_bench.log 2>&1; tail -n 1 gpurun_out/r02ff3_bench.log | cut -c1-300
GIMMVFI_CONV7_SMALL=0 timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02ff3_bench_conv7off.log 2>&1; tail -n 1 gpurun_out/r02ff3_bench_conv7off.log | cut -c1-300
timeout 600 python scripts/f_bench.py --profile-json gpurun_out/r02ff3_f_profile.json > gpurun_out/r02ff3_fbench.log 2>&1; echo "fbench rc=$?"; head -24 gpurun_out/r02ff3_fbench.log | cut -c1-200
timeout 300 python scripts/f_bench.py --h 2176 --w 4096 --ds 0.25 --profile-json gpurun_out/r02ff3_f4k_profile.json > gpurun_out/r02ff3_fbench_4k.log 2>&1; echo "fbench4k rc=$?"; head -8 gpurun_out/r02ff3_fbench_4k.log | cut -c1-300
export GIMMVFI_TC_SPIN_LIMIT=0
for spec in "CostConv1K:0:1:costconv1" "WindowAttnK:0:1:winattn" "layernorm_warp:30:1:layernorm" "conv7x7_small_cout:0:1:conv7" "StridedAttnK<16>:0:1:attn16"; do
  IFS=: read -r pat skip cnt tag <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$pat" -s "$skip" -c "$cnt" -f -o gpurun_out/r02ff3_ncu_$tag \
     python scripts/f_bench.py --steps 1 --warmup 0 > gpurun_out/r02ff3_under_ncu_$tag.log 2>&1; echo "ncu $tag rc=$?"
done
ls -la gpurun_out/r02ff3*.ncu-rep
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python scripts/f_sanitize.py > gpurun_out/r02ff3_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -n 3 gpurun_out/r02ff3_sanitizer_$tool.log | cut -c1-200
done
