#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 --timeout-method=thread > gpurun_out/v5_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/v5_pytest.log; tail -6 gpurun_out/v5_pytest.log
timeout 400 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v5_bench_cfg4.json 2> gpurun_out/v5_bench_cfg4.err
python - <<'Q'
import json
try:
    d=json.loads(open('gpurun_out/v5_bench_cfg4.json').read().strip().splitlines()[-1]); print('cfg4', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
except Exception as e: print('cfg4 ERR', e); print(open('gpurun_out/v5_bench_cfg4.err').read()[-1500:])
Q
timeout 600 python tools/bench_augment.py cfg2 50 > gpurun_out/v5_bench_augment.json 2> gpurun_out/v5_bench_augment.err; cat gpurun_out/v5_bench_augment.json; tail -3 gpurun_out/v5_bench_augment.err
