#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --timeout-method=thread --deselect tests/test_gpu_baseline_configs.py::test_two_rank_nccl_real_plan --durations=8 > gpurun_out/h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/h_pytest.log; tail -14 gpurun_out/h_pytest.log
for wl in cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_$wl.json 2> gpurun_out/h_bench_$wl.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/h_bench_$wl.json').read().strip().splitlines()[-1]); print('$wl', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['launches'])
except Exception as e: print('$wl ERR', e); print(open('gpurun_out/h_bench_$wl.err').read()[-1500:])
P
done
