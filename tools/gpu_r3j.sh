#!/bin/bash
V=$PWD/stormphrax_b200/_lib/variants
echo "== ticket2 racecheck on the slots workload"; SP_NNUE_LIB=$V/slots_ticket2.so timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/prof_slots.py 1 2>&1 | grep -v "Host Frame\|Saved host" | head -60
