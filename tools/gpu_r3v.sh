#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r3v_shuffled_launches.csv python tools/prof_full.py 1048576 1 shuffle > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r3v_shuffled_launches.csv "shuffled" | tail -8
