#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_full.py -m gpu -x -q -k "dense_head or golden_playouts or stress_network" > gpurun_out/r2o_tests_head.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2o_tests_head.log
tail -n 5 gpurun_out/r2o_tests_head.log
if grep -q "rc=0" gpurun_out/r2o_tests_head.log; then
  timeout 300 python tools/head_sweep.py > gpurun_out/r2o_head_sweep_umma.md 2>&1
  SP_NNUE_HEAD=stream timeout 300 python tools/head_sweep.py > gpurun_out/r2o_head_sweep_stream.md 2>&1
  tail -n 16 gpurun_out/r2o_head_sweep_umma.md; tail -n 9 gpurun_out/r2o_head_sweep_stream.md
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2o_tests_all.log
  tail -n 5 gpurun_out/r2o_tests_all.log
fi
