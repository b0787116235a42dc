mkdir -p gpurun_out
for seg in 1 2; do
  echo "=== GIMMVFI_TC_SEG=$seg"
  GIMMVFI_TC_SEG=$seg timeout 900 python scripts/parity_1080p.py 2>&1 | grep -v Warn | grep -E "oracle|mode 2|mode 1" | tee -a gpurun_out/parity_seg.log
done
GIMMVFI_TC_SEG=1 timeout 500 python scripts/tc_e2e_check.py 2>&1 | grep -v Warn | grep -A14 "1080p tc=2" | tee gpurun_out/tc_e2e_seg1.log
timeout 200 python scripts/tc_microbench.py 1 2>&1 | grep -v Warn
