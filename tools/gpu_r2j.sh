#!/bin/bash
# config 5 pieces: the reference's own datagen loop on the GPU evaluator (synchronous evals), and the batched stand-in driver at 5000 nodes per move
mkdir -p gpurun_out /tmp/dg_cpu /tmp/dg_gpu
python -c "
from stormphrax_b200 import net as N
N.synthetic(7, tame=True).image.tofile('/tmp/tame7.nnue')"
( time oracle/_ref/sp_engine_b200 /tmp/tame7.nnue bench 2 ) 2>&1 | grep -E "nodes [0-9]+ nps|^real" | tail -n 2 > gpurun_out/r2j_engine_bench_gpu.log
( time oracle/_ref/sp_engine_cpu /tmp/tame7.nnue bench 2 ) 2>&1 | grep -E "nodes [0-9]+ nps|^real" | tail -n 2 > gpurun_out/r2j_engine_bench_cpu.log
timeout 120 oracle/_ref/sp_engine_b200 /tmp/tame7.nnue datagen /tmp/dg_gpu 40 > gpurun_out/r2j_datagen_gpu.log 2>&1
timeout 120 oracle/_ref/sp_engine_cpu /tmp/tame7.nnue datagen /tmp/dg_cpu 40 > gpurun_out/r2j_datagen_cpu.log 2>&1
ls -la /tmp/dg_gpu /tmp/dg_cpu >> gpurun_out/r2j_datagen_gpu.log
timeout 600 python tools/selfplay_bench.py 65536 1 12 5000 4 1 1 > gpurun_out/r2j_selfplay_5000.json 2> gpurun_out/r2j_selfplay_5000.err
for f in r2j_engine_bench_gpu.log r2j_engine_bench_cpu.log r2j_datagen_gpu.log r2j_datagen_cpu.log r2j_selfplay_5000.json; do echo "== $f"; tail -n 6 gpurun_out/$f; done
