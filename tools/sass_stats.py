#!/usr/bin/env python
"""Static SASS statistics per kernel of libpangaea_b200.so (no GPU needed):
    python tools/sass_stats.py [name-substring]
Prints instruction count and the opcode mix - the fully unrolled 32-window bodies make
'instructions / 32' a fair estimate of the per-window cost before spending GPU time."""
import collections, re, subprocess, sys
LIB = "pangaea_b200/libpangaea_b200.so"
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        funcs[cur][m.group(2)] += 1
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for name, c in funcs.items():
    if pat in name:
        tot = sum(c.values())
        print(f"{name}: {tot} instructions ({tot / 32:.1f} per window if unrolled x32)")
        print("   " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
