#!/bin/bash
mkdir -p gpurun_out
V=$PWD/stormphrax_b200/_lib/variants
timeout 600 python -m pytest tests/test_gpu_full.py -m gpu -x -q -k "dense_head or golden_playouts or stress_network" > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z_tests.log
tail -n 3 gpurun_out/r2z_tests.log
echo "== default (7 gather warps)"; SWEEP_LOGM=14,16,18,20 timeout 300 python tools/head_sweep.py 2>/dev/null | tail -8
for v in umma_notail umma_nogather umma_loaders3 umma_loaders5 umma_notail_loaders3; do
  echo "== $v"; SWEEP_LOGM=20 SP_NNUE_LIB=$V/$v.so timeout 200 python tools/head_sweep.py 2>/dev/null | tail -2
done
