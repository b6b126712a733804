#!/bin/bash
# N-GPU bench lines of round 2: /usr/local/graft/bin/gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi_bench.sh N'
mkdir -p gpurun_out
N=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 600 $RUN 29513 bench.py --gpus $N > gpurun_out/multi_bench_full_${N}gpu.json 2> gpurun_out/multi_bench_full_${N}gpu.err; echo "full rc=$?"
timeout 300 $RUN 29514 bench.py --gpus $N --impl reference > gpurun_out/multi_bench_reference_${N}gpu.json 2> gpurun_out/multi_bench_reference_${N}gpu.err; echo "reference rc=$?"
python - <<PY
import json
for f in ("multi_bench_full_${N}gpu", "multi_bench_reference_${N}gpu"):
    try:
        j = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(j["value"], 1), j["n_gpus"], "e2e", round(j["e2e"]["value"], 1), j.get("parity"))
        for k, v in j.get("workloads", {}).items(): print("  ", k, v.get("value"), v.get("unit"), v.get("parity"))
    except Exception as e:
        print(f, "ERR", e)
PY
