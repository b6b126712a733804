#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_full.py -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
SP_NNUE_LIB=stormphrax_b200/_lib/variants/group_timing.so timeout 120 python tools/prof_full.py 262144 2 > gpurun_out/r2d_timing.log 2>&1
timeout 120 python tools/prof_full.py 1048576 5 > gpurun_out/r2d_rate.log 2>&1
timeout 120 python tools/prof_full.py 1048576 5 shuffle > gpurun_out/r2d_rate_shuffled.log 2>&1
SP_NNUE_FT=warp timeout 120 python tools/prof_full.py 1048576 5 > gpurun_out/r2d_rate_warp.log 2>&1
for f in r2d_tests.log r2d_timing.log r2d_rate.log r2d_rate_shuffled.log r2d_rate_warp.log; do echo "== $f"; tail -n 4 gpurun_out/$f; done
