#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --extras none > gpurun_out/r3i_bench_full.json 2> gpurun_out/r3i_bench_full.err; echo "bench full rc=$?"
SP_NNUE_CHUNK=131072 timeout 600 python bench.py --extras none > gpurun_out/r3i_bench_full_c131072.json 2> /dev/null
SP_NNUE_CHUNK=65536 timeout 600 python bench.py --extras none > gpurun_out/r3i_bench_full_c65536.json 2> /dev/null
python - <<'PY'
import json
for f in ("r3i_bench_full", "r3i_bench_full_c131072", "r3i_bench_full_c65536"):
    try:
        j = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(j["value"], 2), j["unit"], "e2e", round(j["e2e"]["value"], 2), "pageable", round(j["e2e"]["pageable"]["value"], 2), "parity", j.get("parity", {}).get("mismatches"), "roofline launches", j["roofline"]["launches"], j["roofline"]["avg_launch_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
