#!/bin/bash
# GPU call F: ft_full source-level capture + launch list of the full-refresh bench with the streaming head.
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ft_full -c 1 --launch-skip 3 -f -o gpurun_out/ft_full_v11 python tools/kbench.py full 262144 > gpurun_out/ncu_ft_full.log 2>&1; echo "ncu ft_full rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_full_v11.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_full_v11.log 2>&1; echo "launch list rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_v11.json 2> gpurun_out/bench_reference_v11.err; echo "ref arm rc=$?"; cat gpurun_out/bench_reference_v11.json
