"""Per-kernel HBM summary from an `ncu --page raw --csv` export with many launches: for every distinct kernel
(name + grid) keep the launch with the most DRAM bytes and print duration, bytes, GB/s and % of DRAM peak."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except (KeyError, ValueError, IndexError):
        return float("nan")
units = rows[1]
def to_bytes(r, k):
    v, u = num(r, k), units[col[k]] if k in col else ""
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def to_us(r, k):
    v, u = num(r, k), units[col[k]] if k in col else ""
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
def rate_gbs(r):
    k = "dram__bytes.sum.per_second"
    v, u = num(r, k), units[col[k]] if k in col else ""
    return v * {"byte/s": 1e-9, "Kbyte/s": 1e-6, "Mbyte/s": 1e-3, "Gbyte/s": 1, "Tbyte/s": 1e3}.get(u, float("nan"))
def dram_bytes(r):
    if "dram__bytes_read.sum" in col:
        return to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
    return rate_gbs(r) * 1e9 * to_us(r, "gpu__time_duration.sum") * 1e-6
best = {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]].split("(")[0]
    key = (name, r[col["Grid Size"]])
    b = dram_bytes(r)
    if key not in best or b > best[key][0]:
        best[key] = (b, r)
w = csv.writer(sys.stdout)
w.writerow(["kernel", "grid", "block", "regs", "duration_us", "dram_MB", "dram_GB_per_s", "dram_pct_of_peak",
            "lts_pct", "sm_pct", "achieved_occupancy_pct"])
for (name, grid), (b, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
    us = to_us(r, "gpu__time_duration.sum")
    w.writerow([name, grid, r[col["Block Size"]],
                r[col["launch__registers_per_thread"]] if "launch__registers_per_thread" in col else "",
                f"{us:.1f}", f"{b / 1e6:.1f}", f"{b / us / 1e3:.0f}" if us == us and us > 0 else "",
                r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]] if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in col else "",
                r[col["lts__throughput.avg.pct_of_peak_sustained_elapsed"]] if "lts__throughput.avg.pct_of_peak_sustained_elapsed" in col else "",
                r[col["sm__throughput.avg.pct_of_peak_sustained_elapsed"]] if "sm__throughput.avg.pct_of_peak_sustained_elapsed" in col else "",
                r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]] if "sm__warps_active.avg.pct_of_peak_sustained_active" in col else ""])
