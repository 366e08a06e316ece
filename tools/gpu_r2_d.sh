#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:conv_halo_kernel -s 44 -c 1 -o gpurun_out/d_halo32 -f $B > gpurun_out/d_p1.log 2>&1
tail -3 gpurun_out/d_p1.log
ls -la gpurun_out/*.ncu-rep | tail -3
