#!/usr/bin/env bash
# ncu --set full captures of one step per precision mode (every launch of the last iteration), brought back as .ncu-rep
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02n}
for prec in tf32 fp32; do
  OFFK_SINGLE_STREAM=1 timeout -k 5 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -c 130 -f \
    -o $OUT/full_${prec}_$TAG python tools/prof_step.py 48 3 $prec 2 > $OUT/full_${prec}_$TAG.log 2>&1; tail -2 $OUT/full_${prec}_$TAG.log; ls -la $OUT/full_${prec}_$TAG.ncu-rep
  cp $OUT/step_names.txt $OUT/step_names_${prec}_$TAG.txt
done
