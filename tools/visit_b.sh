#!/usr/bin/env bash
# GPU visit B (round 2): parity suite, default bench (fp32 = 3xTF32) with all baselines, launch list
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02b}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout -k 5 240 python tools/x3_probe.py > $OUT/x3_probe_$TAG.log 2>&1; echo "probe exit $?"; grep -E "tf32x3|^M" $OUT/x3_probe_$TAG.log
timeout -k 5 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; grep -E " passed| failed| error" $OUT/pytest_$TAG.log | tail -3
grep -o "\[parity[^]]*\][^[]*" $OUT/pytest_$TAG.log | cut -c1-330
grep -E "^FAILED|^E  +Assertion" $OUT/pytest_$TAG.log | cut -c1-300 | head -40
timeout -k 5 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cut -c1-3000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
OFFK_SINGLE_STREAM=1 timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
   --log-file $OUT/launches_fp32_$TAG.csv python tools/prof_step.py 48 3 fp32 3 > $OUT/launches_fp32_$TAG.log 2>&1
python tools/launch_table.py $OUT/launches_fp32_$TAG.csv $OUT/step_names.txt > $OUT/launches_fp32_$TAG.txt 2>&1; tail -22 $OUT/launches_fp32_$TAG.txt
