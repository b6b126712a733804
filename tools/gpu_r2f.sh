#!/bin/bash
mkdir -p gpurun_out
V=stormphrax_b200/_lib/variants
timeout 300 python tools/prof_slots.py 2>&1 | tail -n 1
SP_NNUE_LIB=$V/slots3.so timeout 300 python tools/prof_slots.py 2>&1 | tail -n 1
SP_NNUE_LIB=$V/slots4.so timeout 300 python tools/prof_slots.py 2>&1 | tail -n 1
