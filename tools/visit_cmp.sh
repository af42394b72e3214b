#!/usr/bin/env bash
# compare unit weight-gradient kernel times: current tree vs the tree in _old/ (bring-up experiment)
set -u
OUT=gpurun_out; mkdir -p $OUT
for tree in . _old; do
  (cd $tree && mkdir -p gpurun_out && OFFK_SINGLE_STREAM=1 timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
     -k regex:tma_gemm_kernel -c 200 --csv --log-file /tmp/l.csv python tools/prof_step.py 48 3 tf32 3 > /tmp/l.log 2>&1; \
   echo "== tree $tree"; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>10 and r[0].isdigit()]
import collections
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0]; grid=r[7] if len(r)>7 else ''
    agg.setdefault(name,[]).append(float(r[-1].replace(',',''))/1000.0)
for k,v in agg.items(): print(f"{k:60s} n={len(v):3d} sum={sum(v):8.1f} us  last7={[round(x,1) for x in v[-9:]]}")
PY
  )
done
