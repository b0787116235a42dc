#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/halo_diag.py > gpurun_out/r02f_halo_diag.log 2>&1; cat gpurun_out/r02f_halo_diag.log | head -60
echo "=== base_offset forced 0"
GIMMVFI_HALO_DBG=1 timeout 200 python scripts/halo_diag.py > gpurun_out/r02f_halo_diag_dbg1.log 2>&1; grep "tap (" gpurun_out/r02f_halo_diag_dbg1.log
