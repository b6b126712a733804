#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2w_bench_full.json 2> gpurun_out/r2w_bench_full.err; echo "bench full rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 1 --extras none > gpurun_out/r2w_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_umma -s 1 -c 1 -f -o gpurun_out/r2w_head_umma python tools/head_once.py 20 3 > gpurun_out/r2w_ncu_head.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ft_group -s 2 -c 1 -f -o gpurun_out/r2w_ft_group python tools/prof_full.py 131072 1 > gpurun_out/r2w_ncu_group.log 2>&1
SWEEP_LOGM=10,11,12 timeout 200 python tools/head_sweep.py 2>/dev/null | tail -6
python -c "
from stormphrax_b200 import net as N
N.synthetic(7, tame=True).image.tofile('/tmp/tame7.nnue')"
( timeout 600 oracle/_ref/sp_engine_b200 /tmp/tame7.nnue bench 2 | tail -n 2 ) 2>&1 | tee gpurun_out/r2w_bench_sync.log
python - <<'PY'
import json
for f in ("r2w_bench_full", "r2w_bench_reference"):
    try:
        j = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(j["value"], 2), j["unit"], "e2e", round(j["e2e"]["value"], 2), "launches", j.get("gpu_launches"), "parity", j.get("parity"))
        w = j.get("workloads", {})
        for k, v in w.items():
            print("  ", k, {kk: vv for kk, vv in v.items() if kk in ("value", "unit", "parity")})
            if k == "head_sweep": print("     ", [(r["M"], round(r["us"], 1), round(r["hbm_frac"], 3)) for r in v["per_gpu"]])
        print("  roofline", j.get("roofline"))
    except Exception as e:
        print(f, "ERR", e)
PY
