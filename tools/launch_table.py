"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-launch table + totals per kernel family.
    python tools/launch_table.py gpurun_out/launches.csv [names.txt]
names.txt (optional): one engine step name per launch, printed beside each row."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
        rows.append((int(r["ID"]), r["Kernel Name"], r["Grid Size"], r["Block Size"], us))
names = open(sys.argv[2]).read().split("\n") if len(sys.argv) > 2 else None
tot = sum(r[4] for r in rows)
fam = defaultdict(lambda: [0, 0.0])
for i, (id_, k, g, b, us) in enumerate(rows):
    short = k.split("(")[0].replace("offk::", "").replace("void ", "")
    fam[short][0] += 1
    fam[short][1] += us
    nm = names[i] if names and i < len(names) else ""
    print(f"{i:4d} {us:9.1f} us  {short[:60]:60s} grid {g:22s} {nm}")
print(f"TOTAL {tot:.1f} us over {len(rows)} launches")
for k, (n, us) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"  {us:9.1f} us {100*us/tot:5.1f}%  x{n:3d}  {k}")
