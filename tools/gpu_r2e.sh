#!/bin/bash
mkdir -p gpurun_out
V=stormphrax_b200/_lib/variants
run() { echo "== $1"; shift; env "$@" timeout 120 python tools/prof_full.py 262144 2 2>&1 | tail -n 3; }
run "timing, 2 CTAs/SM" SP_NNUE_LIB=$V/group_timing.so
run "timing, 1 CTA/SM" SP_NNUE_LIB=$V/group_timing.so SP_NNUE_GROUP_CTAS=1
run "timing unroll2, 2 CTAs/SM" SP_NNUE_LIB=$V/group_timing_unroll2.so
echo "== rates (1M x 5)"
timeout 120 python tools/prof_full.py 1048576 5 | tail -n 1
SP_NNUE_LIB=$V/group_unroll2.so timeout 120 python tools/prof_full.py 1048576 5 | tail -n 1
SP_NNUE_LIB=$V/group_unroll8.so timeout 120 python tools/prof_full.py 1048576 5 | tail -n 1
SP_NNUE_GROUP_CTAS=1 timeout 120 python tools/prof_full.py 1048576 5 | tail -n 1
