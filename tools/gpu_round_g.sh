#!/bin/bash
# GPU call G: one-instruction row addresses (SP_ROW_ADDR) A/B on both FT kernels + full GPU suite.
mkdir -p gpurun_out
V=$PWD/stormphrax_b200/_lib/variants
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/t_all_g.log 2>&1; rc=$?; echo "all gpu tests rc=$rc"; tail -3 gpurun_out/t_all_g.log
if [ $rc -ne 0 ]; then grep -B5 -A25 "Error\|assert" gpurun_out/t_all_g.log | head -60; exit 1; fi
for rep in 1 2; do
  timeout 200 python tools/kbench.py both 2>&1 | tail -2
  SP_NNUE_LIB=$V/rowaddr0.so timeout 200 python tools/kbench.py both 2>&1 | tail -2
done
