#!/bin/bash
# one GPU session: parity suite, bench (with CPU baseline), launch list, full ncu capture of the conv kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1800 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 300 python tools/bench_kernel.py 32 32 64 128 128 2 10 > gpurun_out/kernels.txt 2>&1
timeout 300 python tools/bench_kernel.py 64 32 64 128 128 2 10 >> gpurun_out/kernels.txt 2>&1
timeout 300 python tools/bench_kernel.py 64 64 32 64 64 2 10 >> gpurun_out/kernels.txt 2>&1
timeout 300 python tools/bench_kernel.py 128 128 16 32 32 2 10 >> gpurun_out/kernels.txt 2>&1
timeout 300 python tools/bench_kernel.py 256 256 8 16 16 2 10 >> gpurun_out/kernels.txt 2>&1
timeout 300 python tools/bench_kernel.py 320 320 4 8 8 2 10 >> gpurun_out/kernels.txt 2>&1
cat gpurun_out/kernels.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel" -s 6 -c 3 -o gpurun_out/prof_conv python tools/bench_kernel.py 32 32 64 128 128 2 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
