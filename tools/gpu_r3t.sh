#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_full.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/prof_full.py 1048576 5; timeout 200 python tools/prof_full.py 1048576 5; timeout 200 python tools/prof_full.py 1048576 5 shuffle
timeout 300 python bench.py --extras none 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['parity']['mismatches'])"
