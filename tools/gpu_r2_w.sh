#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_inference.py -m gpu -q --timeout 500 --timeout-method=thread > gpurun_out/w_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/w_pytest.log; tail -30 gpurun_out/w_pytest.log
