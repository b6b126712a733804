#!/bin/bash
# Self-play games with datagen's per-move search (soft node limit) of the reference engine, shared out to host threads:
# CPU engine on T host threads beside the GPU-evaluated engine as T fiber schedulers.  Needs oracle/_ref/tame.nnue
# (stormphrax_b200.net.synthetic(7, tame=True)).  usage: gpu_engine_games.sh <games> <soft nodes> <plies> <threads> <width> [nocpu]
mkdir -p gpurun_out
LOG=gpurun_out/engine_games.log
NET=oracle/_ref/tame.nnue
echo "host cores: $(nproc)" >> $LOG
[ -z "$6" ] && timeout 60 oracle/_ref/sp_engine_cpu $NET games $1 $2 $3 42 1 $4 2>> $LOG | tail -n 1 | sed "s/^/cpu  /" >> $LOG
timeout 80 oracle/_ref/sp_engine_b200 $NET games $1 $2 $3 42 1 $4 $5 2>> $LOG | tail -n 1 | sed "s/^/b200 /" >> $LOG
cat $LOG
