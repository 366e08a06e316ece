#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_augment.py tests/test_gpu_trainers.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/r_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r_pytest.log; tail -30 gpurun_out/r_pytest.log
timeout 600 python tools/bench_augment.py cfg2 50 > gpurun_out/r_bench_augment.json 2> gpurun_out/r_bench_augment.err; cat gpurun_out/r_bench_augment.json; tail -5 gpurun_out/r_bench_augment.err
