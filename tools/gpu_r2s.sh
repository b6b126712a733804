#!/bin/bash
mkdir -p gpurun_out
V=$PWD/stormphrax_b200/_lib/variants
timeout 600 python -m pytest tests/test_gpu_full.py -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s_tests.log
tail -n 3 gpurun_out/r2s_tests.log
echo "== default"; timeout 300 python tools/head_sweep.py 2>/dev/null | tail -22 | tee gpurun_out/r2s_head_sweep.md
for v in umma_notail umma_nogather; do
  echo "== $v"; SWEEP_LOGM=20 SP_NNUE_LIB=$V/$v.so timeout 200 python tools/head_sweep.py 2>/dev/null | tail -2
done
echo "== full refresh"; timeout 200 python tools/prof_full.py 1048576 5; timeout 200 python tools/prof_full.py 1048576 5 shuffle
