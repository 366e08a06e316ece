#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_unet.py tests/test_gpu_fullsize.py tests/test_gpu_trainers.py -m gpu -q -x --timeout 600 --timeout-method=thread > gpurun_out/i_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/i_pytest.log; tail -3 gpurun_out/i_pytest.log
timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q --timeout 600 --timeout-method=thread -k "cfg2 or fused" > gpurun_out/i_pytest2.log 2>&1
echo "pytest exit $?" >> gpurun_out/i_pytest2.log; tail -3 gpurun_out/i_pytest2.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/i_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['with_prefetch_inputs']['value'], 'kernel_ms', d['roofline']['kernel_ms'], d['roofline']['frac'])
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/i_ncu.log 2>&1
