#!/usr/bin/env python
"""Per-kernel shares from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <bench command>).
    python tools/launch_summary.py gpurun_out/launches_r02b.csv [passes]
`passes` = hot-path passes the command ran (default 5: 3 warm-up + 1 timed + 1 for the roofline read-back)."""
import csv, sys, collections

path = sys.argv[1]
passes = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
ms, n = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = r[ix["Kernel Name"]]
    if not name.startswith(("pg::", "void pg::")) or "synth_kernel" in name:
        continue  # torch's own kernels (read-back reductions) and the input generator
    k = name.split("(")[0].replace("void ", "").replace("pg::", "")
    unit = r[ix["Metric Unit"]]
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    ms[k] += v
    n[k] += 1
total = sum(ms.values())
print(f"{'kernel':<42}{'launches/pass':>14}{'ncu ms/pass':>13}{'share':>9}")
for k, v in ms.most_common():
    print(f"{k:<42}{n[k] / passes:>14.1f}{v / passes:>13.3f}{100 * v / total:>8.1f}%")
print(f"\nTotal {total / passes:.1f} ms per pass serialised.")
