#!/bin/bash
# round 2: ncu --set full of the pointwise kernels rewritten last (HBM-bound: achieved GB/s against the measured peak) at the config-3 shapes
mkdir -p gpurun_out
export GIMMVFI_TC_SPIN_LIMIT=0
for spec in "MultiFlowBlendPxK:0:1:blend" "PadZeroK4:40:1:padzero" "conv7x7_small_cout:0:1:conv7v2" "ResizeK4:60:1:resize4"; do
  IFS=: read -r pat skip cnt tag <<< "$spec"
  timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$pat" -s "$skip" -c "$cnt" -f -o gpurun_out/r02last_ncu_$tag \
     python scripts/f_bench.py --steps 1 --warmup 0 > gpurun_out/r02last_under_ncu_$tag.log 2>&1; echo "ncu $tag rc=$?"
done
ls -la gpurun_out/r02last*.ncu-rep
