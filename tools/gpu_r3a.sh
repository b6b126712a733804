#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_full.py -m gpu -x -q > gpurun_out/r3a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3a_tests.log
tail -n 3 gpurun_out/r3a_tests.log
echo "== head sweep"; timeout 300 python tools/head_sweep.py 2>/dev/null | tail -22 | tee gpurun_out/r3a_head_sweep.md
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r3a_head_once_launches.csv python tools/head_once.py 20 3 > /dev/null 2>&1; grep -E "head_" gpurun_out/r3a_head_once_launches.csv | cut -d, -f5,15- | tail -6
echo "== full refresh"; timeout 200 python tools/prof_full.py 1048576 5
