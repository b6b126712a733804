#!/bin/bash
# A/B of the prepared variants (python tools/variants.py first, here, where nvcc is):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_variants.sh'
# Every variant first has to pass the parity tests of the code it touches; only then is it timed.
mkdir -p gpurun_out
V=$PWD/stormphrax_b200/_lib/variants
echo "== default"; SWEEP_LOGM=16,20 timeout 200 python tools/head_sweep.py 2>/dev/null | tail -4; timeout 200 python tools/kbench.py both 2>&1 | tail -2
for v in head_lds head_c12; do
  [ -f $V/$v.so ] || continue
  echo "== $v"
  SP_NNUE_LIB=$V/$v.so timeout 200 python -m pytest tests/test_gpu_full.py -x -q -m gpu -k head > gpurun_out/t_$v.log 2>&1 || { echo "$v FAILED its tests"; tail -5 gpurun_out/t_$v.log; continue; }
  SWEEP_LOGM=16,20 SP_NNUE_LIB=$V/$v.so timeout 200 python tools/head_sweep.py 2>/dev/null | tail -4
done
for v in enq_prefetch enq_unroll2; do
  [ -f $V/$v.so ] || continue
  echo "== $v"
  SP_NNUE_LIB=$V/$v.so timeout 300 python -m pytest tests/test_gpu_full.py tests/test_gpu_incremental.py -x -q -m gpu > gpurun_out/t_$v.log 2>&1 || { echo "$v FAILED its tests"; tail -5 gpurun_out/t_$v.log; continue; }
  SP_NNUE_LIB=$V/$v.so timeout 200 python tools/kbench.py both 2>&1 | tail -2
done
echo "== default self-play"; timeout 200 python tools/selfplay_bench.py 65536 1 2 500 12 1 2 2>/dev/null
for v in slots3 slots4; do
  [ -f $V/$v.so ] || continue
  echo "== $v"
  SP_NNUE_LIB=$V/$v.so timeout 200 python -m pytest tests/test_gpu_incremental.py -x -q -m gpu > gpurun_out/t_$v.log 2>&1 || { echo "$v FAILED its tests"; tail -5 gpurun_out/t_$v.log; continue; }
  SP_NNUE_LIB=$V/$v.so timeout 200 python tools/selfplay_bench.py 65536 1 2 500 12 1 2 2>/dev/null
done
