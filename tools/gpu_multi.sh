#!/bin/bash
# N-GPU call (default 8): GPU-resident self-play (65,536 game slots per GPU) and both bench workloads.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_multi.sh 8'     (ONLY_SELFPLAY=1: skip the bench lines)
mkdir -p gpurun_out
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 300 $RUN 29511 tools/selfplay_bench.py 65536 1 2 500 12 1 2 > gpurun_out/selfplay_resident_${N}gpu.json 2> gpurun_out/selfplay_resident_${N}gpu.err; echo "selfplay rc=$?"; cat gpurun_out/selfplay_resident_${N}gpu.json
[ -n "$ONLY_SELFPLAY" ] && exit 0
timeout 300 $RUN 29512 bench.py --gpus $N --steps 5 --warmup 3 --workload playouts > gpurun_out/bench_playouts_${N}gpu_v11.json 2> gpurun_out/bench_playouts_${N}gpu_v11.err; echo "playouts rc=$?"
timeout 300 $RUN 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_full_${N}gpu_v11.json 2> gpurun_out/bench_full_${N}gpu_v11.err; echo "full rc=$?"
python - <<PY
import json
for f in ("bench_playouts_${N}gpu_v11", "bench_full_${N}gpu_v11"):
    try:
        j = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(j["value"], 1), j["n_gpus"], "e2e", round(j["e2e"]["value"], 1))
    except Exception as e:
        print(f, "ERR", e)
PY
