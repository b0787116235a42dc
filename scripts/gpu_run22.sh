mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_conv_tc_gpu.py -q 2>&1 | tail -3 | tee gpurun_out/pytest_tc.log
timeout 200 python scripts/tc_microbench.py split 2>&1 | grep -v Warn | tee gpurun_out/tc_micro.log
GIMMVFI_TC_SPLIT_EPI8=0 timeout 200 python scripts/tc_microbench.py split 2>&1 | grep -v Warn | sed 's/^/sepi4: /' | tee -a gpurun_out/tc_micro.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --profile-json gpurun_out/profile_bench_default.json 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/bench_default.log
GIMMVFI_TC_SPLIT_EPI8=0 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400 | sed 's/^/sepi4: /' | tee -a gpurun_out/bench_default.log
