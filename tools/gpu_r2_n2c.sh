#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q --timeout 500 --timeout-method=thread -k "two_rank" > gpurun_out/n2c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/n2c_pytest.log; tail -4 gpurun_out/n2c_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n2c_bench.json 2> gpurun_out/n2c_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/n2c_bench.json').read().strip().splitlines()[-1]); print('n2', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))" || tail -8 gpurun_out/n2c_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/n2c_ref.json 2> gpurun_out/n2c_ref.err; tail -c 300 gpurun_out/n2c_ref.json
