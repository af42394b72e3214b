#!/usr/bin/env bash
# deep-K-block TMA weight gradients: parity + per-layer times for bk in {gather, 32, 64, 128}, tf32 and fp32
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02w}
timeout 600 python -m pytest tests -m gpu -q -x -k "tma_gemm_weight_gradient" 2>&1 | tail -4
for prec in tf32 fp32; do
  for cfg in "none 0" "kxk 32" "kxk 64" "kxk 128" "all 64" "all 128"; do
    set -- $cfg
    OFFK_WGRAD_TMA=$1 OFFK_WGRAD_BK=$2 OFFK_SINGLE_STREAM=1 timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
      -k regex:gemm -c 300 --csv --log-file /tmp/l.csv python tools/prof_step.py 48 3 $prec 3 > /tmp/l.log 2>&1
    python - "$prec" "$1" "$2" <<'PY'
import csv, sys
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>10 and r[0].isdigit()]
names=[l.strip() for l in open('gpurun_out/step_names.txt')]
gn=[n for n in names if any(k in n for k in ('unit_','motion_','.wgrad','.dgrad')) and not n.endswith(('.zero','.bias_act')) and 'tapT' not in n and 'relu' not in n]
t=[float(r[-1].replace(',',''))/1000.0 for r in rows]
if len(gn)!=len(t):
    print("name/launch mismatch", len(gn), len(t))
w={n:x for n,x in zip(gn,t) if n.endswith('.wgrad')}
big=[w.get(k+'.wgrad',0) for k in ('motion_conv_trans_28','motion_conv_trans_14','motion_conv_trans')]
small=sum(v for k,v in w.items() if not k.startswith('unit_'))-sum(big)
print(f"{sys.argv[1]} wgrad={sys.argv[2]:5s} bk={sys.argv[3]:4s}: trans_28 {big[0]:6.1f}  trans_14 {big[1]:6.1f}  trans {big[2]:6.1f}  other stage wgrads {small:7.1f}  all gemm {sum(t):7.1f} us")
PY
  done
done
