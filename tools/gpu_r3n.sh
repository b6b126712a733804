#!/bin/bash
for c in 7104 14208 28416 65536; do echo "== SP_NNUE_GAMES_CHUNK=$c"; SP_NNUE_GAMES_CHUNK=$c timeout 300 python bench.py --workload playouts --extras none 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('playouts', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['parity']['mismatches'])"; done
