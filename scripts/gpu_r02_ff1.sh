#!/bin/bash
# round 2, FlowFormer call 1: native GIMM-VFI-F parity on the B200, full GPU suite, first timing / per-kernel profile at BASELINE config 3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_f_gpu.py -q -x -s > gpurun_out/r02ff1_f.log 2>&1; echo "f rc=$?"; grep -E "mode|imgt_pred|passed|failed|Error|error" gpurun_out/r02ff1_f.log | cut -c1-220 | tail -40
timeout 600 python scripts/f_bench.py --profile-json gpurun_out/r02ff1_f_profile.json > gpurun_out/r02ff1_fbench.log 2>&1; echo "fbench rc=$?"; head -45 gpurun_out/r02ff1_fbench.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02ff1_all.log 2>&1; echo "all rc=$?"; tail -n 4 gpurun_out/r02ff1_all.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02ff1_bench.log 2>&1; tail -n 1 gpurun_out/r02ff1_bench.log | cut -c1-400
