#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_incremental.py tests/test_gpu_host_mirror.py tests/test_selfplay.py -m gpu -x -q > gpurun_out/r3b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3b_tests.log
tail -n 3 gpurun_out/r3b_tests.log
timeout 300 python tools/prof_slots.py 5
timeout 300 python tools/prof_slots.py 5
