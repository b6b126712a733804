#!/bin/bash
# GPU call A: head_stream_kernel correctness + A/B against the tiled head + ncu + bench.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_full.py -x -q -m gpu -k "head" > gpurun_out/t_head.log 2>&1; echo "head tests rc=$?"
tail -3 gpurun_out/t_head.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"
tail -3 gpurun_out/t_all.log
timeout 400 python tools/head_sweep.py > gpurun_out/sweep_stream.md 2> gpurun_out/sweep_stream.err; echo "sweep rc=$?"
SWEEP_LOGM=10,14,16,18,20 SP_NNUE_HEAD=tiles timeout 300 python tools/head_sweep.py > gpurun_out/sweep_tiles.md 2>&1
SWEEP_LOGM=10,14,16,18,20 SP_NNUE_LIB=$PWD/stormphrax_b200/_lib/variants/c7.so timeout 300 python tools/head_sweep.py > gpurun_out/sweep_c7.md 2>&1
SWEEP_LOGM=10,14,16,18,20 SP_NNUE_LIB=$PWD/stormphrax_b200/_lib/variants/c11s4.so timeout 300 python tools/head_sweep.py > gpurun_out/sweep_c11s4.md 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_stream -c 1 -f -o gpurun_out/head_stream_v7 python tools/head_once.py 20 2 > gpurun_out/ncu_head.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py > gpurun_out/bench_full_v10.json 2> gpurun_out/bench_full_v10.err; echo "bench rc=$?"
tail -2 gpurun_out/sweep_stream.md gpurun_out/sweep_tiles.md gpurun_out/sweep_c7.md gpurun_out/sweep_c11s4.md
cat gpurun_out/bench_full_v10.json
