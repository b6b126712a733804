#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/dbg_slots.py 2>&1 | head -60 | tee gpurun_out/r3c_memcheck.log
