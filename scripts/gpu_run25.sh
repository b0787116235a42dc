#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/tc_split_probe.py > gpurun_out/split_probe.log 2>&1
cat gpurun_out/split_probe.log
