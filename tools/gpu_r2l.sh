#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_incremental.py tests/test_gpu_host_mirror.py tests/test_engine_dropin.py -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2l_tests.log
python -c "
from stormphrax_b200 import net as N
N.synthetic(7, tame=True).image.tofile('/tmp/tame7.nnue')"
E=oracle/_ref
( timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue bench 2 | tail -n 2 ) > gpurun_out/r2l_bench_sync.log 2>&1
( SP_NNUE_SMALL=0 timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue bench 2 | tail -n 2 ) > gpurun_out/r2l_bench_sync_general.log 2>&1
( timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue searches 512 4 1 | tail -n 1 ) > gpurun_out/r2l_searches_fibers.log 2>&1
( SP_NNUE_SMALL=0 timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue searches 512 4 1 | tail -n 1 ) > gpurun_out/r2l_searches_fibers_general.log 2>&1
( timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue games 1024 5000 2 42 1 | tail -n 1 ) > gpurun_out/r2l_games_fibers.log 2>&1
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py 2>&1 | tail -n 12 ) > gpurun_out/r2l_memcheck.log 2>&1
( SANITIZE_GAMES=4 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py 2>&1 | tail -n 12 ) > gpurun_out/r2l_racecheck.log 2>&1
for f in r2l_tests.log r2l_bench_sync.log r2l_bench_sync_general.log r2l_searches_fibers.log r2l_searches_fibers_general.log r2l_games_fibers.log r2l_memcheck.log r2l_racecheck.log; do echo "== $f"; tail -n 4 gpurun_out/$f | cut -c1-400; done
