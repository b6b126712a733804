#!/bin/bash
mkdir -p gpurun_out
V=$PWD/stormphrax_b200/_lib/variants
for v in umma_notail umma_nogather umma_stages11; do
  echo "== $v"; SWEEP_LOGM=20 SP_NNUE_LIB=$V/$v.so timeout 200 python tools/head_sweep.py 2>/dev/null | tail -2
done
echo "== selfplay test, default"; timeout 300 python -m pytest tests/test_selfplay.py -m gpu -x -q 2>&1 | tail -3
echo "== selfplay test, SP_NNUE_SMALL=0"; SP_NNUE_SMALL=0 timeout 300 python -m pytest tests/test_selfplay.py -m gpu -x -q 2>&1 | tail -3
echo "== selfplay test, SP_NNUE_SMALL_MAPPED=0"; SP_NNUE_SMALL_MAPPED=0 timeout 300 python -m pytest tests/test_selfplay.py -m gpu -x -q 2>&1 | tail -3
echo "== selfplay test, SP_NNUE_HEAD=stream"; SP_NNUE_HEAD=stream timeout 300 python -m pytest tests/test_selfplay.py -m gpu -x -q 2>&1 | tail -3
