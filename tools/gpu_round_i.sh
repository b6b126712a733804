#!/bin/bash
# GPU call I: GPU-resident self-play: parity with the host driver, then throughput.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_selfplay.py -x -q -m gpu > gpurun_out/t_selfplay_i.log 2>&1; rc=$?; echo "selfplay tests rc=$rc"; tail -25 gpurun_out/t_selfplay_i.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 200 python tools/selfplay_bench.py 65536 16 2 500 12 1 > gpurun_out/selfplay_resident_64k.json 2> gpurun_out/selfplay_resident_64k.err; echo "resident rc=$?"; cat gpurun_out/selfplay_resident_64k.json; tail -3 gpurun_out/selfplay_resident_64k.err
timeout 200 python tools/selfplay_bench.py 65536 16 2 500 12 0 > gpurun_out/selfplay_host_64k.json 2> gpurun_out/selfplay_host_64k.err; cat gpurun_out/selfplay_host_64k.json
