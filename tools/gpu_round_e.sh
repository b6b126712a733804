#!/bin/bash
# GPU call E: L2 fast paths (narrow weights / non-negative inputs), single-block sort for small launches.
mkdir -p gpurun_out
V=$PWD/stormphrax_b200/_lib/variants
timeout 200 python -m pytest tests/test_gpu_full.py -x -q -m gpu -k "head" > gpurun_out/t_head_e.log 2>&1; rc=$?; echo "head tests rc=$rc"; tail -3 gpurun_out/t_head_e.log
if [ $rc -ne 0 ]; then grep -B5 -A25 "Error\|assert" gpurun_out/t_head_e.log | head -80; exit 1; fi
timeout 200 python tools/head_sweep.py > gpurun_out/sweep_v9.md 2> gpurun_out/sweep_v9.err; tail -24 gpurun_out/sweep_v9.md
SWEEP_LOGM=16,18,20 SP_NNUE_L2_NARROW=0 timeout 200 python tools/head_sweep.py > gpurun_out/sweep_v9_general.md 2>&1; tail -6 gpurun_out/sweep_v9_general.md
SWEEP_LOGM=16,18,20 SP_NNUE_LIB=$V/c12.so timeout 200 python tools/head_sweep.py > gpurun_out/sweep_v9_c12.md 2>&1; tail -6 gpurun_out/sweep_v9_c12.md
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head_stream -c 1 -f -o gpurun_out/head_stream_v9 python tools/head_once.py 20 2 > gpurun_out/ncu_head_v9.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all_e.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/t_all_e.log
timeout 300 python tools/selfplay_bench.py 65536 16 2 500 12 > gpurun_out/selfplay_64k_v3.json 2> gpurun_out/selfplay_64k_v3.err; cat gpurun_out/selfplay_64k_v3.json
