#!/bin/bash
mkdir -p gpurun_out
V=$PWD/stormphrax_b200/_lib/variants
timeout 300 python -m pytest tests/test_gpu_full.py -m gpu -x -q -k "dense_head or golden_playouts or stress_network" > gpurun_out/r2r_tests_head.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2r_tests_head.log
tail -n 3 gpurun_out/r2r_tests_head.log
echo "== default"; SWEEP_LOGM=16,18,20 timeout 200 python tools/head_sweep.py 2>/dev/null | tail -6
for v in umma_notail umma_nogather umma_notail_p3 umma_notail_s12; do
  echo "== $v"; SWEEP_LOGM=20 SP_NNUE_LIB=$V/$v.so timeout 200 python tools/head_sweep.py 2>/dev/null | tail -2
done
