#!/bin/bash
mkdir -p gpurun_out
V=$PWD/stormphrax_b200/_lib/variants
for v in umma_notail_pf1 umma_notail_pf2 umma_notail_pf4 umma_pf2; do
  echo "== $v"; SWEEP_LOGM=20 SP_NNUE_LIB=$V/$v.so timeout 200 python tools/head_sweep.py 2>/dev/null | tail -2
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_umma -s 1 -c 1 -f -o gpurun_out/r2q_head_umma python tools/head_once.py 20 3 > gpurun_out/r2q_ncu.log 2>&1
tail -n 2 gpurun_out/r2q_ncu.log
ls -la gpurun_out/*.ncu-rep | tail -3
