#!/bin/bash
# round-end evidence run: GPU test-suite, bench line (with CPU baseline), launch list of one step, ncu --set full of the dominant kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1200 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
# dominant kernel: forward halo conv of ctx0.1 = 1st conv_halo launch of a step (11 per step); capture in step 2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 11 -c 1 -o gpurun_out/p_halo32 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/p7.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_halo_kernel -s 5 -c 1 -o gpurun_out/p_wgrad_halo -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/p8.log 2>&1
ls -la gpurun_out/*.ncu-rep
