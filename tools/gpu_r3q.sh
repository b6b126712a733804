#!/bin/bash
V=$PWD/stormphrax_b200/_lib/variants
echo "== default"; timeout 300 python tools/prof_slots.py 5 | tail -1
for v in slots3 slots4; do echo "== $v"; SP_NNUE_LIB=$V/$v.so timeout 300 python tools/prof_slots.py 5 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_full.py tests/test_gpu_incremental.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
