#!/bin/bash
# GPU call J: where does a round of the GPU-resident self-play go?  launch list + one full capture of the step kernel.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 600 --csv --log-file gpurun_out/launches_selfplay_v1.csv python tools/selfplay_bench.py 65536 16 2 500 4 1 > gpurun_out/launches_selfplay_v1.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:selfplay_step --launch-skip 300 -c 1 -f -o gpurun_out/selfplay_step_v1 python tools/selfplay_bench.py 65536 16 2 500 4 1 > gpurun_out/ncu_selfplay_step.log 2>&1; echo "ncu rc=$?"
