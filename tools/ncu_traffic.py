#!/usr/bin/env python
"""profiles/traffic.json from ncu reports: dram__bytes_read.sum + dram__bytes_write.sum per launch of every hot kernel
(bench.py reports the entry of its dominant kernel as roofline.traffic).
    python tools/ncu_traffic.py gpurun_out/prof_X_count.ncu-rep gpurun_out/prof_X_feat.ncu-rep > profiles/traffic.json"""
import csv, io, json, subprocess, sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
NAMES = {"pack_kernel": "pack_kernel", "bucket_scatter_kernel<15, 2>": "bucket_scatter_kernel<15,shared>",
         "bucket_scatter_kernel<15, 1>": "bucket_scatter_kernel<15,feat>", "bucket_scatter_kernel<15, 0>": "bucket_scatter_kernel<15,count>",
         "bucket_split_kernel": "bucket_split_kernel", "sub_apply_kernel": "sub_apply_kernel", "tnf_kernel<4>": "tnf_kernel<4>",
         "bucket_apply_feat_kernel<1>": "bucket_apply_feat_kernel", "bucket_apply_feat_kernel<0>": "bucket_apply_feat_kernel<separate>"}
out = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for d in rows[2:]:
        name = d[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("pg::", "").strip()
        if name not in NAMES or NAMES[name] in out:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(d[ix[m]].replace(",", "")) * UNIT[units[ix[m]]]
        out[NAMES[name]] = int(tot)
json.dump(out, sys.stdout, indent=1)
print()
