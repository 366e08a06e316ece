#!/bin/bash
# full GPU test suite + the four workload benches + the cfg2 launch list (round-2 evidence)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread --durations=8 > gpurun_out/y_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/y_pytest.log; tail -15 gpurun_out/y_pytest.log
for w in cfg2 cfg3 cfg5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/y_bench_$w.json 2> gpurun_out/y_bench_$w.err
done
python - <<'Q'
import json
for w in ('cfg2','cfg3','cfg5'):
    try:
        d=json.loads(open('gpurun_out/y_bench_%s.json'%w).read().strip().splitlines()[-1]); print(w, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'])
    except Exception as e: print(w,'ERR', e); print(open('gpurun_out/y_bench_%s.err'%w).read()[-1500:])
Q
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/y_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/y_ncu.log 2>&1
