#!/usr/bin/env bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full captures (stencil + GEMM).
# usage (under gpurun): bash tools/gpu_round.sh <tag> [what...]   what in {tests bench launches full_stencil full_gemm}
set -u
TAG="${1:-r01}"; shift || true
WHAT="${*:-tests bench launches full_stencil full_gemm}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
for w in $WHAT; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log; tail -5 $OUT/pytest_$TAG.log;;
    sanitize)
      timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k stencil > $OUT/sanitize_$TAG.log 2>&1; echo "sanitize exit $?" >> $OUT/sanitize_$TAG.log; tail -8 $OUT/sanitize_$TAG.log;;
    smoke)
      timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> $OUT/smoke_$TAG.log; tail -3 $OUT/smoke_$TAG.log;;
    issue)
      timeout 600 python tools/issue_time.py 48 3 tf32 > $OUT/issue_$TAG.log 2>&1; timeout 300 python tools/issue_time.py 1 3 tf32 >> $OUT/issue_$TAG.log 2>&1; timeout 300 python tools/issue_time.py 16 7 tf32 >> $OUT/issue_$TAG.log 2>&1; cat $OUT/issue_$TAG.log;;
    bench)
      timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err;;
    benchref)
      timeout 900 python bench.py --impl reference > $OUT/benchref_$TAG.json 2> $OUT/benchref_$TAG.err; cat $OUT/benchref_$TAG.json;;
    launches)
      OFFK_SINGLE_STREAM=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
        --log-file $OUT/launches_$TAG.csv python tools/prof_step.py 48 3 tf32 3 > $OUT/launches_$TAG.log 2>&1
      python tools/launch_table.py $OUT/launches_$TAG.csv $OUT/step_names.txt > $OUT/launches_$TAG.txt 2>&1; tail -25 $OUT/launches_$TAG.txt;;
    full_stencil)
      OFFK_SINGLE_STREAM=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:stencil -c 30 -f -o $OUT/stencil_$TAG python tools/prof_step.py 48 3 tf32 2 > $OUT/full_stencil_$TAG.log 2>&1; tail -2 $OUT/full_stencil_$TAG.log;;
    full_all)
      OFFK_SINGLE_STREAM=1 timeout 1200 ncu --set full --clock-control none --profile-from-start off \
        -c 150 -f -o /tmp/all_$TAG python tools/prof_step.py 48 3 tf32 2 > $OUT/full_all_$TAG.log 2>&1; tail -2 $OUT/full_all_$TAG.log
      ncu -i /tmp/all_$TAG.ncu-rep --page raw --csv > $OUT/full_all_$TAG.csv 2>/dev/null; ls -la /tmp/all_$TAG.ncu-rep;;
    full_gemm)
      OFFK_SINGLE_STREAM=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:tma_gemm_kernel -c 9 -f -o $OUT/gemm_$TAG python tools/prof_step.py 48 3 tf32 2 > $OUT/full_gemm_$TAG.log 2>&1; tail -2 $OUT/full_gemm_$TAG.log;;
  esac
done
ls -la $OUT
