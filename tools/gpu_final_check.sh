#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/final_tests.log
tail -n 3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final_bench_full.json 2> gpurun_out/final_bench_full.err; echo "bench full rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --extras none > gpurun_out/final_bench_under_ncu.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/final_head_launches.csv python tools/head_once.py 20 3 > /dev/null 2>&1
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py 2>&1 | tail -n 4 ) | tee gpurun_out/final_memcheck.log
( SANITIZE_GAMES=4 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py 2>&1 | tail -n 4 ) | tee gpurun_out/final_racecheck.log
python - <<'PY'
import json
for f in ("final_bench_full", "final_bench_reference"):
    try:
        j = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(j["value"], 2), j["unit"], "e2e", round(j["e2e"]["value"], 2), "launches", j.get("gpu_launches"), "parity", j.get("parity"))
        w = j.get("workloads", {})
        for k, v in w.items():
            print("  ", k, {kk: vv for kk, vv in v.items() if kk in ("value", "unit", "parity")})
            if k == "head_sweep": print("     ", [(r["M"], round(r["us"], 1), round(r["hbm_frac"], 3)) for r in v["per_gpu"]])
        r = j.get("roofline") or {}
        print("  roofline", {k: r.get(k) for k in ("kernel", "frac", "avg_launch_ms", "share_of_step")}, "onchip", {k: (r.get("onchip") or {}).get(k) for k in ("frac", "achieved", "peak")})
    except Exception as e:
        print(f, "ERR", e)
PY
