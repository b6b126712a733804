#!/bin/bash
# GPU call B: head_stream_kernel (with the generation gate) + self-play driver.  Bails out early if the new head is wrong.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_full.py -x -q -m gpu -k "head" > gpurun_out/t_head.log 2>&1; rc=$?; echo "head tests rc=$rc"
tail -5 gpurun_out/t_head.log
if [ $rc -ne 0 ]; then
  SP_NNUE_HEAD=tiles timeout 150 python -m pytest tests/test_gpu_full.py -x -q -m gpu -k "head" > gpurun_out/t_head_tiles.log 2>&1; echo "tiles head rc=$?"
  tail -5 gpurun_out/t_head_tiles.log
  exit 1
fi
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"
tail -5 gpurun_out/t_all.log
timeout 300 python tools/head_sweep.py > gpurun_out/sweep_stream.md 2> gpurun_out/sweep_stream.err; echo "sweep rc=$?"
SWEEP_LOGM=10,16,20 SP_NNUE_HEAD=tiles timeout 200 python tools/head_sweep.py > gpurun_out/sweep_tiles.md 2>&1
SWEEP_LOGM=16,20 SP_NNUE_LIB=$PWD/stormphrax_b200/_lib/variants/c7.so timeout 200 python tools/head_sweep.py > gpurun_out/sweep_c7.md 2>&1
SWEEP_LOGM=16,20 SP_NNUE_LIB=$PWD/stormphrax_b200/_lib/variants/c11s4.so timeout 200 python tools/head_sweep.py > gpurun_out/sweep_c11s4.md 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head_stream -c 1 -f -o gpurun_out/head_stream_v7 python tools/head_once.py 20 2 > gpurun_out/ncu_head.log 2>&1; echo "ncu rc=$?"
timeout 200 python tools/selfplay_bench.py 8192 16 2 500 30 > gpurun_out/selfplay_small.json 2> gpurun_out/selfplay_small.err; echo "selfplay rc=$?"
cat gpurun_out/selfplay_small.json
timeout 400 python bench.py > gpurun_out/bench_full_v10.json 2> gpurun_out/bench_full_v10.err; echo "bench rc=$?"
tail -2 gpurun_out/sweep_stream.md gpurun_out/sweep_tiles.md gpurun_out/sweep_c7.md gpurun_out/sweep_c11s4.md
cat gpurun_out/bench_full_v10.json
