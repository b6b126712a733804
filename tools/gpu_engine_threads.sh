#!/bin/bash
# The reference engine's own search on the GPU evaluator, shared out to host threads (one fiber scheduler + evaluator context each),
# beside the CPU engine on the same host threads.  usage: gpu_engine_threads.sh <searches> "<threads>:<width> ..." [cpu threads]
# Output: gpurun_out/engine_threads.log
mkdir -p gpurun_out
LOG=gpurun_out/engine_threads.log
python - <<'PY'
from stormphrax_b200 import net as N
N.synthetic(7, tame=True).image.tofile('/tmp/tame.nnue')
PY
echo "host cores: $(nproc)" >> $LOG
N=${1:-2048}
CPU_T=${3:-16}
timeout 90 oracle/_ref/sp_engine_cpu /tmp/tame.nnue searches $N 4 1 $CPU_T > /tmp/cpu.out 2>> $LOG; tail -n 1 /tmp/cpu.out | sed "s/^/cpu  /" >> $LOG
for tw in ${2:-16:64}; do
  t=${tw%%:*}; w=${tw##*:}
  timeout 90 oracle/_ref/sp_engine_b200 /tmp/tame.nnue searches $N 4 1 $t $w > /tmp/b200.out 2>> $LOG; tail -n 1 /tmp/b200.out | sed "s/^/b200 /" >> $LOG
  if [ "$(grep '^search nodes:' /tmp/b200.out | md5sum)" == "$(grep '^search nodes:' /tmp/cpu.out | md5sum)" ]; then echo "     node count of every search identical to the CPU engine's" >> $LOG; else echo "     NODE COUNTS DIFFER (or a run was cut off)" >> $LOG; fi
done
cat $LOG
