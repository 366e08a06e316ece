#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vit_native.py tests/test_gpu_vit_unet.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/u_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/u_pytest.log; tail -60 gpurun_out/u_pytest.log
