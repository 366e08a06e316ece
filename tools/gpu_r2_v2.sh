#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vit_unet.py tests/test_gpu_vit_native.py -m gpu -q --timeout 500 --timeout-method=thread > gpurun_out/v2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/v2_pytest.log; tail -6 gpurun_out/v2_pytest.log
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q --timeout 500 --timeout-method=thread -k cfg4 > gpurun_out/v2_pytest2.log 2>&1
echo "pytest exit $?" >> gpurun_out/v2_pytest2.log; tail -6 gpurun_out/v2_pytest2.log
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v2_bench_cfg4.json 2> gpurun_out/v2_bench_cfg4.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/v2_bench_cfg4.json').read().strip().splitlines()[-1]); print('cfg4', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['launches'])
except Exception as e: print('cfg4 ERR', e); print(open('gpurun_out/v2_bench_cfg4.err').read()[-1500:])
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/v2_launches_cfg4.csv python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/v2_ncu.log 2>&1
