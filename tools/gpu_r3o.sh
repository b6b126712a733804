#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_umma -s 1 -c 1 -f -o gpurun_out/r3o_head_umma python tools/head_once.py 20 3 > gpurun_out/r3o_ncu_head.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ft_group -s 1 -c 1 -f -o gpurun_out/r3o_ft_group python tools/prof_full.py 262144 1 > gpurun_out/r3o_ncu_group.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ft_slots -s 20 -c 1 -f -o gpurun_out/r3o_ft_slots python tools/prof_slots.py 1 > gpurun_out/r3o_ncu_slots.log 2>&1
ls -la gpurun_out/r3o_*.ncu-rep
