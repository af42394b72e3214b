# ncu --set full of single tma_gemm_kernel launches picked by their index among the tma_gemm_kernel launches of a step
# usage: bash tools/prof_two.sh name:skip [name:skip ...]
OUT=gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; skip=${spec##*:}
  OFFK_SINGLE_STREAM=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tma_gemm_kernel --launch-skip $skip -c 1 -f -o $OUT/$name python tools/prof_step.py 48 3 tf32 2 > $OUT/$name.log 2>&1; tail -1 $OUT/$name.log
done
ls -la $OUT/*.ncu-rep
