#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vit_native.py tests/test_gpu_vit_unet.py -m gpu -q --timeout 500 --timeout-method=thread > gpurun_out/x_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/x_pytest.log; tail -15 gpurun_out/x_pytest.log
timeout 400 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/x_bench_cfg4.json 2> gpurun_out/x_bench_cfg4.err
python - <<'Q'
import json
try:
    d=json.loads(open('gpurun_out/x_bench_cfg4.json').read().strip().splitlines()[-1]); print('cfg4', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
except Exception as e: print('cfg4 ERR', e); print(open('gpurun_out/x_bench_cfg4.err').read()[-1500:])
Q
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x_launches_cfg4.csv python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/x_ncu.log 2>&1
