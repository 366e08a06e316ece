#!/bin/bash
# N = 2 scaling experiments: bucket size and NCCL channel count (SM share of the collective)
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/n2b_$name.json 2> gpurun_out/n2b_$name.err
  python - "$name" <<'Q'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/n2b_%s.json'%n).read().strip().splitlines()[-1]); print(n, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
except Exception as e: print(n,'ERR',e); print(open('gpurun_out/n2b_%s.err'%n).read()[-800:])
Q
}
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/n2b_n1.json 2> gpurun_out/n2b_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/n2b_n1.json').read().strip().splitlines()[-1]); print('n1', round(d['value'],1), round(d['ms_per_step'],3))"
run default A=1
run onebucket B2_BUCKET_MB=100000
run b16 B2_BUCKET_MB=16
run b8 B2_BUCKET_MB=8
run ch4 NCCL_MAX_NCHANNELS=4
run ch8 NCCL_MAX_NCHANNELS=8
run ch8b16 NCCL_MAX_NCHANNELS=8 B2_BUCKET_MB=16
