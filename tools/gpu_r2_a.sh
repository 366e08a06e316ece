#!/bin/bash
# round 2, run A: new parity tests at the BASELINE configs (all results, no -x), the rest of the GPU suite, smoke, bench cfg2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_env.txt; nproc >> gpurun_out/a_env.txt; free -g >> gpurun_out/a_env.txt
timeout 1500 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q --timeout 900 --timeout-method=thread --deselect tests/test_gpu_baseline_configs.py::test_two_rank_nccl_real_plan > gpurun_out/a_pytest_new.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest_new.log
tail -5 gpurun_out/a_pytest_new.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method=thread --ignore tests/test_gpu_baseline_configs.py > gpurun_out/a_pytest_old.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest_old.log
tail -5 gpurun_out/a_pytest_old.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; tail -3 gpurun_out/a_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -c 1500 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/a_bench_nograph.json 2> gpurun_out/a_bench_nograph.err
tail -c 600 gpurun_out/a_bench_nograph.json; tail -3 gpurun_out/a_bench_nograph.err
