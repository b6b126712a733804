#!/bin/bash
mkdir -p gpurun_out
SP_NNUE_LIB=stormphrax_b200/_lib/variants/group_timing.so timeout 120 python tools/prof_full.py 262144 2 > gpurun_out/r2c_timing.log 2>&1
SP_NNUE_LIB=stormphrax_b200/_lib/variants/group_timing.so timeout 120 python tools/prof_full.py 262144 2 shuffle > gpurun_out/r2c_timing_shuffled.log 2>&1
timeout 120 python tools/prof_full.py 1048576 5 > gpurun_out/r2c_rate.log 2>&1
timeout 120 python tools/prof_full.py 1048576 5 shuffle > gpurun_out/r2c_rate_shuffled.log 2>&1
SP_NNUE_FT=warp timeout 120 python tools/prof_full.py 1048576 5 shuffle > gpurun_out/r2c_rate_shuffled_warp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ft_group -s 4 -c 1 -o gpurun_out/r2c_ft_group python tools/prof_full.py 131072 1 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_timing.log gpurun_out/r2c_timing_shuffled.log gpurun_out/r2c_rate.log gpurun_out/r2c_rate_shuffled.log gpurun_out/r2c_rate_shuffled_warp.log
tail -2 gpurun_out/r2c_ncu.log
