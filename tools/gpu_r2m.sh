#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_incremental.py tests/test_gpu_host_mirror.py tests/test_engine_dropin.py -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.log
python -c "
from stormphrax_b200 import net as N
N.synthetic(7, tame=True).image.tofile('/tmp/tame7.nnue')"
E=oracle/_ref
( timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue bench 2 | tail -n 2 ) > gpurun_out/r2m_bench_sync.log 2>&1
( SP_NNUE_SMALL_MAPPED=0 timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue bench 2 | tail -n 2 ) > gpurun_out/r2m_bench_sync_copies.log 2>&1
( timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue searches 512 4 1 | tail -n 1 ) > gpurun_out/r2m_searches_fibers.log 2>&1
( SP_NNUE_SMALL_MAPPED=0 timeout 600 $E/sp_engine_b200 /tmp/tame7.nnue searches 512 4 1 | tail -n 1 ) > gpurun_out/r2m_searches_fibers_copies.log 2>&1
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py 2>&1 | tail -n 12 ) > gpurun_out/r2m_memcheck.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu 2> /dev/null
timeout 300 /tmp/ubench gpurun_out/onchip_peaks.json > gpurun_out/r2m_ubench.log 2>&1
for f in r2m_tests.log r2m_bench_sync.log r2m_bench_sync_copies.log r2m_searches_fibers.log r2m_searches_fibers_copies.log r2m_memcheck.log r2m_ubench.log; do echo "== $f"; tail -n 6 gpurun_out/$f | cut -c1-400; done
