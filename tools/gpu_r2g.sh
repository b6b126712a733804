#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ft_slots -s 20 -c 1 -f -o gpurun_out/r2g_ft_slots python tools/prof_slots.py 1 > gpurun_out/r2g_ncu_slots.log 2>&1
tail -n 3 gpurun_out/r2g_ncu_slots.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ft_group -s 2 -c 1 -f -o gpurun_out/r2g_ft_group python tools/prof_full.py 131072 1 > gpurun_out/r2g_ncu_group.log 2>&1
tail -n 8 gpurun_out/r2g_ncu_group.log
ls -la gpurun_out/*.ncu-rep
