#!/bin/bash
# full GPU suite + default bench + reference arm on the current tree
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2h_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?" >> gpurun_out/r2h_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err
tail -n 3 gpurun_out/r2h_tests.log; tail -n 2 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2h_bench.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "pageable", round(d["e2e"]["pageable"]["value"], 1), "parity", d["parity"])
for k, v in d.get("workloads", {}).items():
    if k == "head_sweep":
        print(k, [(r["M"], round(r["us"], 1), round(r["hbm_frac"], 3)) for r in v["per_gpu"]])
    else:
        print(k, round(v["value"], 1), v["unit"], "e2e", round(v["e2e"]["value"], 1), "parity", v["parity"])
r = json.load(open("gpurun_out/r2h_bench_reference.json"))
print("reference", round(r["value"], 3), r["cpu_baseline"]["cores"], "cores;", r.get("workloads"))
PY
