#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_env.txt
timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q --timeout 500 --timeout-method=thread -k "two_rank" > gpurun_out/n2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/n2_pytest.log; tail -15 gpurun_out/n2_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
tail -c 400 gpurun_out/n2_bench.json; tail -5 gpurun_out/n2_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-graph > gpurun_out/n2_bench_nograph.json 2> gpurun_out/n2_bench_nograph.err
tail -c 300 gpurun_out/n2_bench_nograph.json; tail -5 gpurun_out/n2_bench_nograph.err
python - <<'P'
import json
for f in ['gpurun_out/n2_bench.json','gpurun_out/n2_bench_nograph.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
P
