#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel", 34), ("gpu__time_duration.sum", "ms", 8), ("dram__bytes_read.sum", "rd", 8), ("dram__bytes_write.sum", "wr", 8),
        ("lts__t_sector_hit_rate.pct", "L2hit", 6), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%", 6),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1%", 6), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM%", 6),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 7), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 6),
        ("launch__registers_per_thread", "regs", 5), ("smsp__inst_executed.sum", "winst", 10),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 6)]
print(" ".join(f"{n:>{w}}" for _, n, w in cols))
for d in data:
    out = []
    for h, n, w in cols:
        v = d[idx[h]] if h in idx else ""
        if h == "Kernel Name":
            v = v.split("(")[0].replace("void ", "").replace("pg::", "")[:w]
        else:
            try:
                f = float(v.replace(",", ""))
                u = units[idx[h]]
                if n == "winst": v = f"{f/1e6:.1f}M"
                elif n in ("rd", "wr"): v = f"{f:.3f}{u[0] if u else ''}"
                else: v = f"{f:.2f}"
            except ValueError:
                pass
        out.append(f"{v:>{w}}")
    print(" ".join(out))
if len(sys.argv) > 2:  # stall reasons
    want = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") or ("warp_issue_stalled" in h and h.endswith(".pct"))]
    for d in data:
        print(d[idx["Kernel Name"]].split("(")[0])
        vals = sorted(((float(d[idx[h]] or 0), h) for h in want), reverse=True)[:8]
        for v, h in vals: print(f"    {v:8.2f} {h}")
