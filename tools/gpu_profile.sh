#!/bin/bash
# targeted `ncu --set full` captures (one launch each) of the kernels under investigation; read here with
#   ncu -i gpurun_out/<name>.ncu-rep --page details   (or --page raw --csv / --page source --csv)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
N="ncu --set full --clock-control none --import-source on"
# conv_tc launches per step: 42 (first-layer GEMM = #0, tconv tu.4 forward = #21, strided dgrad ctx1.0 = #41); capture step 2
timeout 300 $N -k regex:conv_tc_kernel -s 42 -c 1 -o gpurun_out/p_first_gemm -f $B > gpurun_out/p1.log 2>&1
timeout 300 $N -k regex:conv_tc_kernel -s 63 -c 1 -o gpurun_out/p_tconv_fwd -f $B > gpurun_out/p2.log 2>&1
timeout 300 $N -k regex:conv_tc_kernel -s 83 -c 1 -o gpurun_out/p_dgrad_strided -f $B > gpurun_out/p3.log 2>&1
timeout 300 $N -k regex:wgrad_reduce_tiled_kernel -s 18 -c 18 -o gpurun_out/p_wgrad_reduce -f $B > gpurun_out/p4.log 2>&1
timeout 300 $N -k regex:"norm_bwd_finalize|stats_finalize" -s 24 -c 4 -o gpurun_out/p_finalize -f $B > gpurun_out/p5.log 2>&1
timeout 300 $N -k regex:wgrad_tc_kernel -s 27 -c 2 -o gpurun_out/p_wgrad_tc -f $B > gpurun_out/p6.log 2>&1
timeout 300 $N -k regex:conv_halo_kernel -s 11 -c 1 -o gpurun_out/p_halo32 -f $B > gpurun_out/p7.log 2>&1
ls -la gpurun_out/*.ncu-rep
