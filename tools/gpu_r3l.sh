#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_full.py -m gpu -x -q -k "dense_head or device_pointer or counters" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r3l_bench_full.json 2> gpurun_out/r3l_bench_full.err; echo "bench full rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r3l_bench_full.json").read().strip().splitlines()[-1])
print(round(j["value"], 2), "e2e", round(j["e2e"]["value"], 2), "parity", j["parity"]["mismatches"])
for k, v in j["workloads"].items():
    print("  ", k, v.get("value"), v.get("parity"))
    if k == "head_sweep": print("     ", [(r["M"], round(r["us"], 1), round(r["hbm_frac"], 3), round(r.get("kernel_us", 0), 1), round(r.get("kernel_hbm_frac", 0), 3)) for r in v["per_gpu"]])
PY
