#!/usr/bin/env bash
# same-box sweep of one environment variable: tools/visit_env.sh <tag> <VAR> <prec> v1 v2 ...   (two interleaved passes)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=$1; VAR=$2; PREC=$3; shift 3
rm -f $OUT/env_$TAG.log
for rep in 1 2; do
  for v in "$@"; do
    echo "== rep $rep $VAR=$v" >> $OUT/env_$TAG.log
    env $VAR=$v timeout -k 5 300 python tools/issue_time.py 48 3 $PREC 2>&1 | grep -E "GRAPH replay both" >> $OUT/env_$TAG.log
  done
done
cat $OUT/env_$TAG.log
