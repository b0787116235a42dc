#!/bin/bash
# round 2, call n: what does a tile of the 3xF16 split kernel cost besides its K loop?  (K sweep of the GRU gate shape, stall counters, epilogue variants)
mkdir -p gpurun_out
PROBE_GRU=1 timeout 900 python scripts/tc_split_probe.py PROBE_STALL=1 PROBE_STALL=1,PROBE_EPI=tanh PROBE_STALL=1,PROBE_EPI=q PROBE_STALL=1,GIMMVFI_TC_SEG_F16=4 PROBE_STALL=1,GIMMVFI_TC_PAIR=0 > gpurun_out/r02n_gru_probe.log 2>&1
cat gpurun_out/r02n_gru_probe.log | cut -c1-1500
