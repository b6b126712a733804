#!/bin/bash
# GPU call D: head with straight-line elected copy issue (default: self-refill, 8 warps); full GPU suite; benches.
mkdir -p gpurun_out
V=$PWD/stormphrax_b200/_lib/variants
timeout 120 python -m pytest tests/test_gpu_full.py -x -q -m gpu -k "head" > gpurun_out/t_head_d.log 2>&1; rc=$?; echo "head tests rc=$rc"; tail -2 gpurun_out/t_head_d.log
if [ $rc -ne 0 ]; then exit 1; fi
for v in c12 c6; do
  SP_NNUE_LIB=$V/$v.so timeout 120 python -m pytest tests/test_gpu_full.py -x -q -m gpu -k "head" > gpurun_out/t_head_$v.log 2>&1; echo "$v head tests rc=$?"
done
timeout 200 python tools/head_sweep.py > gpurun_out/sweep_v8.md 2> gpurun_out/sweep_v8.err; tail -6 gpurun_out/sweep_v8.md
for v in c12 c6; do
  SWEEP_LOGM=16,18,20 SP_NNUE_LIB=$V/$v.so timeout 200 python tools/head_sweep.py > gpurun_out/sweep_v8_$v.md 2>&1; tail -6 gpurun_out/sweep_v8_$v.md
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head_stream -c 1 -f -o gpurun_out/head_stream_v8 python tools/head_once.py 20 2 > gpurun_out/ncu_head_v8.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all_d.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/t_all_d.log
timeout 400 python bench.py > gpurun_out/bench_full_v11.json 2> gpurun_out/bench_full_v11.err; echo "bench rc=$?"
timeout 400 python bench.py --workload playouts > gpurun_out/bench_playouts_v11.json 2> gpurun_out/bench_playouts_v11.err; echo "bench playouts rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_full_v11.json","gpurun_out/bench_playouts_v11.json"):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j["value"], j["e2e"]["value"], j["roofline"].get("other_kernels_ms_per_launch"), j["roofline"]["avg_launch_ms"])
    except Exception as e: print(f, "ERR", e)
PY
