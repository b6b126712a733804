#!/bin/bash
# One GPU call that re-checks the whole repository: GPU test suite, smoke(), both bench workloads, the reference arm.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full rc=$?"
timeout 400 python bench.py --workload playouts > gpurun_out/bench_playouts.json 2> gpurun_out/bench_playouts.err; echo "bench playouts rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"
python - <<'PY'
import json
for f in ("bench_full", "bench_playouts", "bench_reference"):
    try:
        j = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(j["value"], 2), j["unit"], "e2e", round(j["e2e"]["value"], 2), "launches", j.get("gpu_launches"))
    except Exception as e:
        print(f, "ERR", e)
PY
