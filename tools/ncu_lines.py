#!/usr/bin/env python
"""Per CUDA source line: executed warp instructions and stall samples from an .ncu-rep (needs -lineinfo + --import-source)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kernel = sys.argv[3] if len(sys.argv) > 3 else None  # optional kernel-name regex
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kernel:
    cmd += ["--kernel-name", f"regex:{kernel}"]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
agg = {}
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0]:
        d = dict(zip(hdr, r))
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1]])
        a[0] += int(d.get("Instructions Executed") or 0)
        a[1] += int(d.get("# Samples") or 0)
tot_i = sum(a[0] for a in agg.values())
tot_s = sum(a[1] for a in agg.values())
print(f"total warp instructions {tot_i}, samples {tot_s}")
for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{f}:{line:4d} inst {100*a[0]/max(tot_i,1):5.1f}%  samples {100*a[1]/max(tot_s,1):5.1f}%  {a[2].strip()[:90]}")
