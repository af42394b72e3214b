"""One line per launch from the metrics CSV that tools/visit_ncu.sh brings back (ncu --page raw --csv, filtered):
    python tools/ncu_table.py gpurun_out/ncu_<tag>_<prec>_metrics.csv [step_names.txt]
time, DRAM rate, L2 throughput and hit rate, shared-memory bank traffic, tensor-pipe activity, occupancy, registers, shared
memory, grid and waves.  (Reads the CSV, not the .ncu-rep: the reports stay on the GPU box.)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
names = [l.strip() for l in open(sys.argv[2])] if len(sys.argv) > 2 else []
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def g(r, k, default=float("nan")):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return default


def dram_tbs(r):
    v, u = g(r, "dram__bytes.sum.per_second"), units[col["dram__bytes.sum.per_second"]] if "dram__bytes.sum.per_second" in col else ""
    # ncu picks the unit per column, not per row: convert by the stated unit
    return v * {"Tbyte/s": 1.0, "Gbyte/s": 1e-3, "Mbyte/s": 1e-6, "byte/s": 1e-12}.get(u, 1.0)


print("# " + " ".join(sys.argv[1:]))
print(f"{'#':>3} {'us':>7} {'dramTB/s':>8} {'L2%':>5} {'L2hit':>5} {'smemRd%':>7} {'smemWr%':>7} {'tensor%':>7} {'occ%':>5} {'regs':>4} {'smemKB':>6} {'grid':>14} {'waves':>5}  kernel / step")
tot = 0.0
for i, r in enumerate(rows[2:]):
    if not r or len(r) < len(hdr) // 2:
        continue
    t = g(r, "gpu__time_duration.sum")
    tot += t
    kn = r[col["Kernel Name"]].split("(")[0].replace("void ", "")[:40]
    grid = f"({int(g(r, 'launch__grid_dim_x', 0))},{int(g(r, 'launch__grid_dim_y', 0))},{int(g(r, 'launch__grid_dim_z', 0))})"
    nm = names[i] if i < len(names) else ""
    print(f"{i:3d} {t:7.1f} {dram_tbs(r):8.2f} {g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} {g(r, 'lts__t_sector_hit_rate.pct'):5.1f} "
          f"{g(r, 'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed'):7.1f} {g(r, 'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed'):7.1f} "
          f"{g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f} {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} "
          f"{g(r, 'launch__registers_per_thread'):4.0f} {g(r, 'launch__shared_mem_per_block'):6.1f} {grid:>14} {g(r, 'launch__waves_per_multiprocessor'):5.2f}  {kn}  {nm}")
print(f"TOTAL {tot:.1f} us over {len(rows) - 2} launches")
