#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainers.py tests/test_gpu_augment.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/z2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/z2_pytest.log; tail -25 gpurun_out/z2_pytest.log
timeout 900 python bench.py > gpurun_out/z2_bench_cfg2.json 2> gpurun_out/z2_bench_cfg2.err
for w in cfg3 cfg5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/z2_bench_$w.json 2> gpurun_out/z2_bench_$w.err
done
python - <<'Q'
import json
for w in ('cfg2','cfg3','cfg5'):
    try:
        d=json.loads(open('gpurun_out/z2_bench_%s.json'%w).read().strip().splitlines()[-1]); print(w, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(w,'ERR', e); print(open('gpurun_out/z2_bench_%s.err'%w).read()[-1500:])
Q
