#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3r_tests.log
tail -n 3 gpurun_out/r3r_tests.log
timeout 300 python tools/prof_slots.py 5 | tail -1
( SANITIZE_GAMES=4 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py 2>&1 | tail -n 3 )
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py 2>&1 | tail -n 3 )
