#!/usr/bin/env bash
# A/B of library builds on ONE box (box-to-box variance is ~1-2 %): tools/visit_ab.sh <tag> <lib_a> <lib_b> ...
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-ab}; shift
rm -f $OUT/ab_$TAG.log
for rep in 1 2; do
  for lib in "$@"; do
    for shape in "48 3 tf32" "48 3 fp32"; do
      echo "== rep $rep lib=$lib  ${OFFK_AB_ENV:-}" >> $OUT/ab_$TAG.log
      OFFK_LIB=$PWD/$lib timeout -k 5 300 python tools/issue_time.py $shape 2>&1 | grep -E "GRAPH replay both" >> $OUT/ab_$TAG.log
    done
  done
done
cat $OUT/ab_$TAG.log
