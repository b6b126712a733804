#!/bin/bash
# GPU call C: self-refilling head variants (no producer warp), fused EvalBatch flush, self-play tests + bench.
mkdir -p gpurun_out
V=$PWD/stormphrax_b200/_lib/variants
for v in p0c12 p0c8 p0c16; do
  SP_NNUE_LIB=$V/$v.so timeout 120 python -m pytest tests/test_gpu_full.py -x -q -m gpu -k "head" > gpurun_out/t_head_$v.log 2>&1; echo "$v head tests rc=$?"
  tail -2 gpurun_out/t_head_$v.log
done
timeout 600 python -m pytest tests/test_selfplay.py tests/test_gpu_host_mirror.py tests/test_gpu_incremental.py -x -q -m gpu > gpurun_out/t_selfplay.log 2>&1; echo "selfplay+mirror tests rc=$?"
tail -15 gpurun_out/t_selfplay.log
for v in p0c12 p0c8 p0c16; do
  SWEEP_LOGM=14,16,18,20 SP_NNUE_LIB=$V/$v.so timeout 200 python tools/head_sweep.py > gpurun_out/sweep_$v.md 2>&1
  tail -8 gpurun_out/sweep_$v.md
done
timeout 200 python tools/selfplay_bench.py 8192 16 2 500 30 > gpurun_out/selfplay_small2.json 2> gpurun_out/selfplay_small2.err; echo "selfplay rc=$?"
cat gpurun_out/selfplay_small2.json
timeout 400 python tools/selfplay_bench.py 65536 16 2 500 12 > gpurun_out/selfplay_64k.json 2> gpurun_out/selfplay_64k.err; echo "selfplay 64k rc=$?"
cat gpurun_out/selfplay_64k.json
