#!/bin/bash
# GPU call H: occupancy variants of the walker (3 / 4 CTAs per SM) and the full refresh (3 / 5 CTAs per SM).
V=$PWD/stormphrax_b200/_lib/variants
timeout 200 python tools/kbench.py both 2>&1 | tail -2
SP_NNUE_GAMES_CHUNK=10656 SP_NNUE_LIB=$V/g3.so timeout 200 python tools/kbench.py playouts 2>&1 | tail -1
SP_NNUE_LIB=$V/g3.so timeout 200 python tools/kbench.py playouts 2>&1 | tail -1
SP_NNUE_GAMES_CHUNK=14208 SP_NNUE_LIB=$V/g4.so timeout 200 python tools/kbench.py playouts 2>&1 | tail -1
SP_NNUE_LIB=$V/f3.so timeout 200 python tools/kbench.py full 2>&1 | tail -1
SP_NNUE_LIB=$V/f5.so timeout 200 python tools/kbench.py full 2>&1 | tail -1
