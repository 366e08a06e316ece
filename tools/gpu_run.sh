#!/bin/bash
# one GPU session: parity suite, then bench + launch list when green enough
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
rc=$?
echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
tail -c 1300 gpurun_out/bench_tc.json; tail -3 gpurun_out/bench_tc.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
