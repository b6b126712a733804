#!/bin/bash
V=$PWD/stormphrax_b200/_lib/variants
echo "== ticket2 plain"; SP_NNUE_LIB=$V/slots_ticket2.so timeout 100 python tools/dbg_slots.py 2>&1 | tail -2
echo "== ticket2 racecheck"; SP_NNUE_LIB=$V/slots_ticket2.so timeout 600 compute-sanitizer --tool racecheck python tools/dbg_slots.py 2>&1 | grep -v "^=========     \(Host\|Saved\)" | head -40
echo "== ticket2 memcheck"; SP_NNUE_LIB=$V/slots_ticket2.so timeout 600 compute-sanitizer --tool memcheck python tools/dbg_slots.py 2>&1 | grep -v "Host Frame" | head -30
