#!/bin/bash
V=$PWD/stormphrax_b200/_lib/variants
echo "== default (ticket 2)"; timeout 200 python tools/dbg_slots.py 2>&1 | tail -2
echo "== static"; SP_NNUE_LIB=$V/slots_static.so timeout 200 python tools/dbg_slots.py 2>&1 | tail -2
echo "== ticket 1"; SP_NNUE_LIB=$V/slots_ticket1.so timeout 200 python tools/dbg_slots.py 2>&1 | tail -2
