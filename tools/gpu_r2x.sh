#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_full.py -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2x_tests.log
tail -n 3 gpurun_out/r2x_tests.log
echo "== full refresh"; timeout 200 python tools/prof_full.py 1048576 5; timeout 200 python tools/prof_full.py 1048576 5; timeout 200 python tools/prof_full.py 1048576 5 shuffle
