#!/bin/bash
mkdir -p gpurun_out

timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1100 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
rm -f gpurun_out/kernels.txt
for shp in "32 32 64 128 128" "64 32 64 128 128" "64 64 32 64 64" "128 64 32 64 64" "128 128 16 32 32" "256 256 8 16 16" "320 320 4 8 8" "640 320 4 8 8"; do
  timeout 300 python tools/bench_kernel.py $shp 2 10 >> gpurun_out/kernels.txt 2>&1
done
cat gpurun_out/kernels.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
