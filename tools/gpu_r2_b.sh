#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q --timeout 900 --timeout-method=thread --deselect tests/test_gpu_baseline_configs.py::test_two_rank_nccl_real_plan --durations=15 > gpurun_out/b_pytest_new.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest_new.log
tail -30 gpurun_out/b_pytest_new.log
