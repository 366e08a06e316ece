#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p_aug_launches.csv python tools/profile_augment.py > gpurun_out/p_aug.log 2>&1
python tools/summarize_launches.py gpurun_out/p_aug_launches.csv | head -20
