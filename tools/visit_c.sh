#!/usr/bin/env bash
# GPU visit C (round 2): CUDA graphs + fused training step + device seeds; issue-time table; bench with graphs
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02c}
timeout -k 5 900 python -m pytest tests -m gpu -q -s  > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; grep -E " passed| failed| error" $OUT/pytest_$TAG.log | tail -3
grep -E "^FAILED|^E  +|Error" $OUT/pytest_$TAG.log | cut -c1-300 | head -40
for shape in "1 3" "48 3" "16 7"; do timeout -k 5 300 python tools/issue_time.py $shape tf32 >> $OUT/issue_$TAG.log 2>&1; done; grep -E "both|GRAPH" $OUT/issue_$TAG.log
timeout -k 5 300 python tools/issue_time.py 1 3 fp32 2>&1 | grep -E "both|GRAPH" | tee -a $OUT/issue_$TAG.log
timeout -k 5 600 python bench.py --no-cpu-baseline --no-gpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cut -c1-700 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
timeout -k 5 600 python bench.py --no-cpu-baseline --no-gpu-baseline --no-graphs --no-families > $OUT/bench_nograph_$TAG.json 2> $OUT/bench_nograph_$TAG.err; echo "bench(no graphs) exit $?"; cut -c1-400 $OUT/bench_nograph_$TAG.json
