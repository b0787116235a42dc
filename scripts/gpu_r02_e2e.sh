#!/bin/bash
# end-to-end timing stability of bench.py (longer untimed warm-up of the pipelined loop): three runs of the default line without the baselines
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/r02e2e_bench_$i.log 2>&1; python - gpurun_out/r02e2e_bench_$i.log <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('dev ms %.2f'%l['ms_per_step'], 'e2e ms %.2f'%l['e2e']['ms_per_step'], l['clocks'])
PY
done
timeout 300 python bench.py --no-cpu-baseline --no-torch-baseline --steps 10 > gpurun_out/r02e2e_bench_10.log 2>&1; tail -n 1 gpurun_out/r02e2e_bench_10.log | cut -c1-200
