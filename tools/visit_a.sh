#!/usr/bin/env bash
# GPU visit A (round 2): 3xTF32 bring-up -- probe, parity tests, both bench modes, launch list of the fp32 mode
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout -k 5 240 python tools/x3_probe.py > $OUT/x3_probe_$TAG.log 2>&1; echo "probe exit $?"; cat $OUT/x3_probe_$TAG.log | tail -20
timeout -k 5 1200 python -m pytest tests -m gpu -q -s -k "not benchmark_shapes" > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|error" $OUT/pytest_$TAG.log | tail -5
grep -E "^\[parity|FAILED" $OUT/pytest_$TAG.log | tail -60
timeout -k 5 900 python -m pytest tests -m gpu -q -s -k "benchmark_shapes" > $OUT/pytest_shapes_$TAG.log 2>&1; echo "pytest shapes exit $?"; grep -E "^\[parity|FAILED|passed|failed" $OUT/pytest_shapes_$TAG.log | tail -20
timeout -k 5 400 python bench.py --precision fp32 --no-cpu-baseline > $OUT/bench_fp32_$TAG.json 2> $OUT/bench_fp32_$TAG.err; echo "bench fp32 exit $?"; cut -c1-600 $OUT/bench_fp32_$TAG.json
timeout -k 5 400 python bench.py --precision tf32 --no-cpu-baseline > $OUT/bench_tf32_$TAG.json 2> $OUT/bench_tf32_$TAG.err; echo "bench tf32 exit $?"; cut -c1-600 $OUT/bench_tf32_$TAG.json
OFFK_SINGLE_STREAM=1 timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
   --log-file $OUT/launches_fp32_$TAG.csv python tools/prof_step.py 48 3 fp32 3 > $OUT/launches_fp32_$TAG.log 2>&1
python tools/launch_table.py $OUT/launches_fp32_$TAG.csv $OUT/step_names.txt > $OUT/launches_fp32_$TAG.txt 2>&1; tail -25 $OUT/launches_fp32_$TAG.txt
