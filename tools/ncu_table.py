"""One line per launch of an `ncu --set full` report: python tools/ncu_table.py rep.ncu-rep [names.txt]
time, DRAM bytes and % of peak, L2 (lts) throughput %, L1/shared (l1tex) throughput %, tensor pipe % active, occupancy, registers."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
names = [l.strip() for l in open(sys.argv[2])] if len(sys.argv) > 2 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def g(r, k, scale=1.0, default=float("nan")):
    try:
        return float(r[col[k]].replace(",", "")) * scale
    except Exception:
        return default


print(f"{'#':>3} {'us':>7} {'dram_rd_MB':>10} {'dram_wr_MB':>10} {'dram%':>6} {'L2%':>6} {'L2hit%':>6} {'l1tex%':>6} {'tensor%':>7} {'occ%':>5} {'regs':>4}  kernel / step")
tot = 0.0
for i, r in enumerate(rows[2:]):
    t = g(r, "gpu__time_duration.sum", 1e-3)
    tot += t
    kn = r[col["Kernel Name"]].split("(")[0][:44]
    nm = names[i] if i < len(names) else ""
    print(f"{i:3d} {t:7.1f} {g(r, 'dram__bytes_read.sum', 1e-6):10.1f} {g(r, 'dram__bytes_write.sum', 1e-6):10.1f} "
          f"{g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{g(r, 'lts__t_sector_hit_rate.pct'):6.1f} {g(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f} {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} "
          f"{g(r, 'launch__registers_per_thread'):4.0f}  {kn}  {nm}")
print(f"TOTAL {tot:.1f} us over {len(rows) - 2} launches (units row: {rows[1][col['gpu__time_duration.sum']]})")
