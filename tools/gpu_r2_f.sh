#!/bin/bash
mkdir -p gpurun_out
MASKS=0,16,2 timeout 100 python tools/probe_halo.py 32 32 > gpurun_out/f_probe32c.txt 2>&1; cat gpurun_out/f_probe32c.txt
MASKS=0 timeout 100 python tools/probe_halo.py 64 32 > gpurun_out/f_probe64c.txt 2>&1; cat gpurun_out/f_probe64c.txt
MASKS=0 timeout 100 python tools/probe_halo.py 64 64 32 64 64 > gpurun_out/f_probe6464c.txt 2>&1; cat gpurun_out/f_probe6464c.txt
