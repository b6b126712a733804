#!/bin/bash
V=$PWD/stormphrax_b200/_lib/variants
echo "== ticket2 + syncwarp: parity"; SP_NNUE_LIB=$V/slots_ticket2.so timeout 300 python tools/prof_slots.py 3 2>&1 | tail -1
echo "== ticket2 + syncwarp: racecheck"; SP_NNUE_LIB=$V/slots_ticket2.so timeout 900 compute-sanitizer --tool racecheck python tools/prof_slots.py 1 2>&1 | grep -v "Host Frame\|Saved host" | grep -E "hazard|RACECHECK|slots:" | sort | uniq -c | sort -rn | head -12
echo "== default (ticket 1): racecheck"; timeout 900 compute-sanitizer --tool racecheck python tools/prof_slots.py 1 2>&1 | grep -v "Host Frame\|Saved host" | grep -E "hazard|RACECHECK|slots:" | sort | uniq -c | sort -rn | head -12
