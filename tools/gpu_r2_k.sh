#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_augment.py tests/test_gpu_bf16.py tests/test_gpu_unet.py -m gpu -q -x --timeout 600 --timeout-method=thread > gpurun_out/k_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/k_pytest.log; tail -25 gpurun_out/k_pytest.log
for v in 0 1; do
B2_OPTIONS="splitk_fuse=$v" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/k_bench_$v.json 2> gpurun_out/k_bench_$v.err
python -c "
import json; d=json.loads(open('gpurun_out/k_bench_$v.json').read().strip().splitlines()[-1]); print('splitk_fuse=$v', round(d['value'],1), round(d['ms_per_step'],3), 'launches', d['launches']['per_step'])" || tail -5 gpurun_out/k_bench_$v.err
done
