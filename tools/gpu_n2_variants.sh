#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/n2_$name.err | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', 'value %.1f (%.2f ms)  e2e %.1f (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
}
run default A=1
run maxconn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run noloss_streams B2_LOSS_STREAMS=0
run nooverlap B2_OPTIONS=bwd_overlap=0 B2_LOSS_STREAMS=0
