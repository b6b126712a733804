#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_dropin.py -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2k_tests.log
python -c "
from stormphrax_b200 import net as N
N.synthetic(7, tame=True).image.tofile('/tmp/tame7.nnue')"
E=oracle/_ref
( timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue games 1024 5000 2 42 1 | tail -n 1 ) > gpurun_out/r2k_games_fibers.log 2>&1
( timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue games 32 5000 2 42 0 | tail -n 1 ) > gpurun_out/r2k_games_sync.log 2>&1
( timeout 600 $E/sp_engine_cpu /tmp/tame7.nnue games 32 5000 2 42 0 | tail -n 1 ) > gpurun_out/r2k_games_cpu.log 2>&1
( timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue searches 512 4 1 | tail -n 1 ) > gpurun_out/r2k_searches_fibers.log 2>&1
for f in r2k_tests.log r2k_games_fibers.log r2k_games_sync.log r2k_games_cpu.log r2k_searches_fibers.log; do echo "== $f"; tail -n 3 gpurun_out/$f | cut -c1-400; done
