#!/bin/bash
# round 2: GIMM-VFI-F under CUDA-graph replay, full GPU suite of the final code
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_f_gpu.py -q -x -k "graph or native" > gpurun_out/r02fin2_f.log 2>&1; echo "f rc=$?"; tail -n 3 gpurun_out/r02fin2_f.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02fin2_all.log 2>&1; echo "all rc=$?"; tail -n 3 gpurun_out/r02fin2_all.log | cut -c1-200
