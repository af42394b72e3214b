#!/usr/bin/env bash
# launch list (ncu per-launch durations, serialised) of one fwd+bwd step: tools/visit_l.sh <tag> <prec> [B L]
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02l}; PREC=${2:-tf32}; B=${3:-48}; LG=${4:-3}
OFFK_SINGLE_STREAM=1 timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
   --log-file $OUT/launches_${PREC}_$TAG.csv python tools/prof_step.py $B $LG $PREC 3 > $OUT/launches_${PREC}_$TAG.log 2>&1
python tools/launch_table.py $OUT/launches_${PREC}_$TAG.csv $OUT/step_names.txt > $OUT/launches_${PREC}_$TAG.txt 2>&1; cat $OUT/launches_${PREC}_$TAG.txt
