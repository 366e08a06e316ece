#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q --timeout 900 --timeout-method=thread --deselect tests/test_gpu_baseline_configs.py::test_two_rank_nccl_real_plan -k "cfg1 or cfg4 or cfg5 or mib" > gpurun_out/c_pytest_new.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest_new.log
tail -8 gpurun_out/c_pytest_new.log
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:"conv_halo_kernel<32>" -s 10 -c 1 -o gpurun_out/c_halo32 -f $B > gpurun_out/c_p1.log 2>&1
timeout 300 $N -k regex:"wgrad_halo_kernel" -s 10 -c 1 -o gpurun_out/c_wgrad_halo -f $B > gpurun_out/c_p2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_launches.csv $B > gpurun_out/c_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
