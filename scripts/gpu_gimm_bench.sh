#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gimm_bench.py > gpurun_out/gimm_bench.log 2>&1; cat gpurun_out/gimm_bench.log
