#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02g}
bash tools/visit_l.sh ${TAG}a tf32 > /dev/null 2>&1; grep -E "splitk|head|TOTAL|x +[0-9]+ +" $OUT/launches_tf32_${TAG}a.txt | head -60
echo "=== OFFK_FINISH_MIN_TILES=100"
OFFK_FINISH_MIN_TILES=100 bash tools/visit_l.sh ${TAG}b tf32 > /dev/null 2>&1; grep -E "splitk|bias_act|sum_14b|TOTAL" $OUT/launches_tf32_${TAG}b.txt | head -40
timeout 300 python -m pytest tests -m gpu -q -x -k "finisher or fused_head or fp32_mode or training" 2>&1 | tail -3
