#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_augment.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/q_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/q_pytest.log; tail -5 gpurun_out/q_pytest.log
timeout 600 python tools/bench_augment.py cfg2 50 > gpurun_out/q_bench_augment.json 2> gpurun_out/q_bench_augment.err; cat gpurun_out/q_bench_augment.json; tail -5 gpurun_out/q_bench_augment.err
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
N="ncu --set full --clock-control none --import-source on"
timeout 400 $N -k regex:conv_halo_kernel -s 36 -c 12 -o gpurun_out/q_halo -f $B > gpurun_out/q_p1.log 2>&1
timeout 300 $N -k regex:wgrad_halo_kernel -s 15 -c 5 -o gpurun_out/q_wgrad_halo -f $B > gpurun_out/q_p2.log 2>&1
ls -la gpurun_out/q_*.ncu-rep
