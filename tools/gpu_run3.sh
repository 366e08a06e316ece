#!/bin/bash
# GPU session 3: parity suite (non-TC first), TC conv tests, bench + profiles if TC is green
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --deselect tests/test_gpu_conv.py -k "not bf16_tc" > gpurun_out/pytest_main.log 2>&1
echo "pytest_main exit $?" >> gpurun_out/pytest_main.log
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 120 --timeout-method=thread > gpurun_out/pytest_conv.log 2>&1
rc=$?
echo "pytest_conv exit $rc" >> gpurun_out/pytest_conv.log
tail -5 gpurun_out/pytest_main.log; tail -8 gpurun_out/pytest_conv.log
if [ $rc -eq 0 ]; then
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
  tail -c 1200 gpurun_out/bench_tc.json; tail -3 gpurun_out/bench_tc.err
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
fi
