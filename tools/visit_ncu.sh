#!/usr/bin/env bash
# ncu captures of one step per precision mode (every launch of the last iteration): summary tables come back, the reports stay
# on the box (two full reports are ~400 MB).  One `--set full` report of the dominant kernels is kept per mode (no sources).
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02n}
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats"
for prec in tf32 fp32; do
  OFFK_SINGLE_STREAM=1 timeout -k 5 900 ncu $SECT --clock-control none --profile-from-start off -c 130 -f \
    -o /tmp/all_${prec} python tools/prof_step.py 48 3 $prec 2 > $OUT/ncu_${prec}_$TAG.log 2>&1; tail -1 $OUT/ncu_${prec}_$TAG.log
  cp $OUT/step_names.txt $OUT/step_names_${prec}_$TAG.txt
  ncu -i /tmp/all_${prec}.ncu-rep --page raw --csv > /tmp/all_${prec}.csv 2>/dev/null; python - $prec $TAG <<'PY'
import csv, sys
prec, tag = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(f"/tmp/all_{prec}.csv")))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if any(k in h for k in ("Kernel Name", "gpu__time_duration", "dram__bytes", "dram__throughput", "lts__t_bytes", "lts__t_sector_hit",
        "lts__throughput", "l1tex__throughput", "l1tex__data_pipe", "l1tex__data_bank", "sm__pipe_tensor", "sm__inst_executed_pipe_uniform", "sm__throughput",
        "sm__warps_active", "launch__", "smsp__average_warps_issue_stalled", "smsp__issue_active", "smsp__inst_executed.sum", "sm__cycles_elapsed", "shared"))]
with open(f"gpurun_out/ncu_{tag}_{prec}_metrics.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in keep if i < len(r)])
PY
  python tools/ncu_table.py $OUT/ncu_${TAG}_${prec}_metrics.csv $OUT/step_names_${prec}_$TAG.txt > $OUT/ncu_${TAG}_${prec}_table.txt 2>&1; tail -2 $OUT/ncu_${TAG}_${prec}_table.txt
  ls -la $OUT/ncu_${TAG}_${prec}_*; rm -f /tmp/all_${prec}.ncu-rep
done
