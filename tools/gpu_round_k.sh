#!/bin/bash
# GPU call K: pin-aware movegen + concurrent resident instances.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_selfplay.py -x -q -m gpu > gpurun_out/t_selfplay_k.log 2>&1; rc=$?; echo "selfplay tests rc=$rc"; tail -5 gpurun_out/t_selfplay_k.log
if [ $rc -ne 0 ]; then grep -B5 -A30 "Error" gpurun_out/t_selfplay_k.log | head -80; exit 1; fi
for inst in 1 2 4; do
  timeout 200 python tools/selfplay_bench.py 65536 $inst 2 500 12 1 > gpurun_out/selfplay_resident_64k_i$inst.json 2> gpurun_out/selfplay_resident_i$inst.err; cat gpurun_out/selfplay_resident_64k_i$inst.json
done
timeout 200 python tools/selfplay_bench.py 65536 16 2 500 12 0 > gpurun_out/selfplay_host_64k_v5.json 2>/dev/null; cat gpurun_out/selfplay_host_64k_v5.json
