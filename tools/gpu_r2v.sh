#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2v_tests.log
tail -n 4 gpurun_out/r2v_tests.log
echo "== head sweep (default)"; timeout 300 python tools/head_sweep.py 2>/dev/null | tail -22 | tee gpurun_out/r2v_head_sweep.md
echo "== head sweep, direct off"; SWEEP_LOGM=10,12,13 SP_NNUE_HEAD_DIRECT=0 timeout 300 python tools/head_sweep.py 2>/dev/null | tail -6
python -c "
from stormphrax_b200 import net as N
N.synthetic(7, tame=True).image.tofile('/tmp/tame7.nnue')"
E=oracle/_ref
( timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue bench 2 | tail -n 2 ) 2>&1 | tee gpurun_out/r2v_bench_sync.log
( timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue searches 512 4 1 | tail -n 1 ) 2>&1 | tee gpurun_out/r2v_searches_fibers.log
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py 2>&1 | tail -n 6 ) | tee gpurun_out/r2v_memcheck.log
( SANITIZE_GAMES=4 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py 2>&1 | tail -n 6 ) | tee gpurun_out/r2v_racecheck.log
