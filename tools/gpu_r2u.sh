#!/bin/bash
mkdir -p gpurun_out
V=$PWD/stormphrax_b200/_lib/variants
for v in umma_k4 umma_notail_k4 umma_notail_k2p4; do
  echo "== $v"; SP_NNUE_LIB=$V/$v.so timeout 100 python -m pytest tests/test_gpu_full.py -m gpu -x -q -k "dense_head_large" 2>&1 | tail -1
  SWEEP_LOGM=20 SP_NNUE_LIB=$V/$v.so timeout 200 python tools/head_sweep.py 2>/dev/null | tail -2
done
