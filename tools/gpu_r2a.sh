#!/bin/bash
# Round 2, first GPU call: layout probe for the tcgen05 kernels, on-chip bandwidths, tests, bench, sanitizer.
mkdir -p gpurun_out
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvidia-smi -L > gpurun_out/r2a_gpu.txt 2>&1
nvcc $ARCH -O3 -o /tmp/umma_probe tools/umma_probe.cu 2> /dev/null
for v in 0 1; do timeout 30 /tmp/umma_probe $v; echo "rc=$?"; done > gpurun_out/r2a_umma_probe.log 2>&1
nvcc $ARCH -O3 -o /tmp/ubench tools/ubench.cu 2> /dev/null
timeout 120 /tmp/ubench gpurun_out/onchip_peaks.json > gpurun_out/r2a_ubench.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r2a_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r2a_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2a_racecheck.log
tail -3 gpurun_out/r2a_umma_probe.log gpurun_out/r2a_tests.log gpurun_out/r2a_memcheck.log gpurun_out/r2a_racecheck.log
cat gpurun_out/r2a_ubench.log
