#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vit_native.py -m gpu -q --timeout 500 --timeout-method=thread > gpurun_out/v_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/v_pytest.log; tail -40 gpurun_out/v_pytest.log
