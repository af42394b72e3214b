#!/usr/bin/env bash
# deep-K-block TMA weight gradients: parity + per-layer times for bk in {gather, 32, 64, 128}, tf32 and fp32
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02w}; LOG=$OUT/wgrad_sweep_$TAG.txt; : > $LOG
timeout 600 python -m pytest tests -m gpu -q -k "tma_gemm_weight_gradient" 2>&1 | tail -15 | cut -c1-250 | tee -a $LOG
for prec in tf32 fp32; do
  for cfg in "none 0" "big 128" "kxk 64" "all 64" "all 32"; do
    set -- $cfg
    OFFK_WGRAD_TMA=$1 OFFK_WGRAD_BK=$2 OFFK_SINGLE_STREAM=1 timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
      -c 300 --csv --log-file /tmp/l.csv python tools/prof_step.py 48 3 $prec 3 > /tmp/l.log 2>&1 || tail -5 /tmp/l.log | cut -c1-300 | tee -a $LOG
    python - "$prec" "$1" "$2" <<'PY' | tee -a $LOG
import csv, sys
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>10 and r[0].isdigit()]
names=[l.strip() for l in open('gpurun_out/step_names.txt')]
t=[float(r[-1].replace(',',''))/1000.0 for r in rows]
if len(names)!=len(t):
    print("name/launch mismatch", len(names), len(t))
w={n:x for n,x in zip(names,t) if n.endswith('.wgrad')}
big=[w.get(k+'.wgrad',0) for k in ('motion_conv_trans_28','motion_conv_trans_14','motion_conv_trans')]
small=sum(v for k,v in w.items() if not k.startswith('unit_') and not k.startswith('head'))-sum(big)
print(f"{sys.argv[1]} wgrad={sys.argv[2]:5s} bk={sys.argv[3]:4s}: trans_28 {big[0]:6.1f}  trans_14 {big[1]:6.1f}  trans {big[2]:6.1f}  other stage wgrads {small:7.1f}  step (serialised) {sum(t):7.1f} us")
if sys.argv[2] in ('kxk','all'):
    print("     per layer:", " ".join(f"{k.replace('motion_','').replace('.wgrad','')}={v:.0f}" for k,v in w.items() if not k.startswith(('unit_','head'))))
PY
  done
done
