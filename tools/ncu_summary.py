#!/usr/bin/env python
"""Summarise an .ncu-rep: per kernel duration, DRAM bytes, DRAM throughput %, occupancy, top stalls."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
def g(r, name, default=""):
    i = col.get(name); return r[i] if i is not None and i < len(r) else default
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
stalls = [n for n in hdr if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio")]
if not stalls:
    stalls = [n for n in hdr if n.startswith("smsp__average_warp_latency_issue_stalled") or ("issue_stalled" in n and n.endswith(".ratio"))]
def num(s):
    try: return float(s.replace(",", ""))
    except Exception: return float("nan")
print(f"{'kernel':60s} {'us':>9s} {'rd MB':>9s} {'wr MB':>9s} {'dram%':>6s} {'sm%':>6s} {'occ%':>6s} {'regs':>5s} {'grid':>8s}")
for r in data:
    name = g(r, "Kernel Name")[:58]
    t = num(g(r, "gpu__time_duration.sum")); tu = units[col["gpu__time_duration.sum"]]
    t_us = t / 1e3 if tu in ("ns", "nsecond") else (t if tu.startswith("u") else t * 1e3)
    def mb(name):
        v = num(g(r, name)); u = units[col[name]] if name in col else ""
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return v * mult / 1e6
    print(f"{name:60s} {t_us:9.1f} {mb('dram__bytes_read.sum'):9.1f} {mb('dram__bytes_write.sum'):9.1f} "
          f"{num(g(r, want[3])):6.1f} {num(g(r, want[4])):6.1f} {num(g(r, want[5])):6.1f} {g(r, want[6]):>5s} {g(r, want[7]):>8s}")
    top = sorted(((num(g(r, s)), s) for s in stalls), reverse=True)[:3]
    print("      stalls: " + ", ".join(f"{s.split('issue_stalled_')[-1].replace('_per_issue_active.ratio','')}={v:.2f}" for v, s in top if v == v))
