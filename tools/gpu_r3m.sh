#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full.py tests/test_gpu_incremental.py -m gpu -x -q > gpurun_out/r3m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3m_tests.log
tail -n 3 gpurun_out/r3m_tests.log
timeout 300 python bench.py --workload playouts --extras none 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('playouts', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['parity'])"
