#!/bin/bash
# Round 2, second GPU call: the tensor-core group kernel (parity, timing A/B against the warp kernel), the reference
# engine on the GPU evaluator.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_full.py -m gpu -x -q -k "golden_playouts or golden_dfrc" > gpurun_out/r2b_first.log 2>&1
echo "first rc=$?" >> gpurun_out/r2b_first.log
tail -5 gpurun_out/r2b_first.log
if grep -q "first rc=0" gpurun_out/r2b_first.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
  timeout 600 python bench.py --steps 10 --warmup 3 --extras none > gpurun_out/r2b_bench_group.json 2> gpurun_out/r2b_bench_group.err
  SP_NNUE_FT=warp timeout 600 python bench.py --steps 10 --warmup 3 --extras none > gpurun_out/r2b_bench_warp.json 2> gpurun_out/r2b_bench_warp.err
  python - <<'PY'
import json
for k in ("group", "warp"):
    try:
        d = json.load(open(f"gpurun_out/r2b_bench_{k}.json"))
        print(k, round(d["value"], 1), "Mpos/s  e2e", round(d["e2e"]["value"], 1), "parity", d["parity"]["mismatches"], "ft ms/launch", round(d["roofline"]["avg_launch_ms"], 3), d["roofline"]["other_kernels_ms_per_launch"])
    except Exception as e:
        print(k, "failed:", e)
PY
else
  SP_NNUE_FT=warp timeout 300 python -m pytest tests/test_gpu_full.py -m gpu -x -q -k "golden_playouts" 2>&1 | tail -3
fi
tail -4 gpurun_out/r2b_tests.log
