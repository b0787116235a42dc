#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tc_f16_probe.py > gpurun_out/f16_probe.log 2>&1; cat gpurun_out/f16_probe.log
GIMMVFI_TC_EPI8_WIDE=0 timeout 300 python scripts/tc_f16_probe.py > gpurun_out/f16_probe_ew4.log 2>&1; cat gpurun_out/f16_probe_ew4.log
