#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_incremental.py tests/test_gpu_host_mirror.py tests/test_selfplay.py tests/test_engine_dropin.py -m gpu -x -q > gpurun_out/r3e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3e_tests.log
tail -n 3 gpurun_out/r3e_tests.log
timeout 300 python tools/prof_slots.py 5
SP_NNUE_LIB=$PWD/stormphrax_b200/_lib/variants/slots_static.so timeout 300 python tools/prof_slots.py 5
( SANITIZE_GAMES=4 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py 2>&1 | tail -n 4 )
