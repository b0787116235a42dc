mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
nproc | tee -a gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -rA 2>&1 | tail -120 > gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --profile-json gpurun_out/profile_1080p.json 2>&1 | tail -3 | tee gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2d_simt_kernel -s 300 -c 3 -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out
