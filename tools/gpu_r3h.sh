#!/bin/bash
for c in 65536 131072 262144 524288; do echo "== SP_NNUE_CHUNK=$c"; SP_NNUE_CHUNK=$c timeout 200 python tools/prof_full.py 1048576 5; done
