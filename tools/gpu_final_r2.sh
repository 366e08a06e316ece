#!/bin/bash
# round-2 closing evidence: full GPU suite, smoke(), the four workload benches (+ CPU baseline on the default one), launch lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --timeout-method=thread --durations=6 > gpurun_out/f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/f_pytest.log; tail -12 gpurun_out/f_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/f_smoke.log; tail -3 gpurun_out/f_smoke.log
timeout 900 python bench.py > gpurun_out/f_bench_cfg2.json 2> gpurun_out/f_bench_cfg2.err
for w in cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_$w.json 2> gpurun_out/f_bench_$w.err
done
python - <<'Q'
import json
for w in ('cfg2','cfg3','cfg4','cfg5'):
    try:
        d=json.loads(open('gpurun_out/f_bench_%s.json'%w).read().strip().splitlines()[-1]); print(w, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(w,'ERR', e); print(open('gpurun_out/f_bench_%s.err'%w).read()[-1500:])
Q
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/f_ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches_cfg4.csv python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/f_ncu4.log 2>&1
timeout 300 python tools/bench_augment.py cfg2 50 > gpurun_out/f_bench_augment.json 2> gpurun_out/f_bench_augment.err; cat gpurun_out/f_bench_augment.json
