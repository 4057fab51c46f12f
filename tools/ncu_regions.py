#!/usr/bin/env python
"""Dynamic instruction counts of one kernel by SOURCE REGION of its top-level .cu file.

Joins (a) `ncu -i rep --page source --csv --print-source sass` (executed warp instructions and stall samples per SASS
instruction) with (b) `nvdisasm -gi` of the same cubin (line info incl. inline chains): every SASS instruction is
attributed to the outermost line of the kernel's own file, so inlined helpers count where they are called.

  tools/ncu_regions.py rep.ncu-rep cubin mangled-substring file.cu [bucket-size]
"""
import collections
import csv
import re
import subprocess
import sys

rep, cubin, sym, top = sys.argv[1:5]
import os
kfilter = ["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []  # reports with several kernels
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", *kfilter], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
# a report with several kernels prints one table per kernel: NCU_SECTION=n picks the n-th (default: all rows)
section = os.environ.get("NCU_SECTION", "0" if kfilter else None)
if section is not None:
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"] + [len(rows)]
    rows = rows[starts[int(section)]:starts[int(section) + 1]]
dyn = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
inside, notes, static = False, [], []
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = sym in ln
        notes = []
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        notes.append(m.groups())
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        line = None
        for f, l, f2, l2 in notes:  # the outermost frame that lies in the kernel's own file
            if f2 and f2.endswith(top):
                line = int(l2)
            elif not f2 and f.endswith(top):
                line = int(l)
        inner = notes[0][:2] if notes else ("?", "0")
        static.append((line, inner[0].split("/")[-1], int(inner[1]), ln.split("*/", 1)[1].strip()[:60]))
        notes_keep = notes
        notes = []
        if line is None and static[:-1]:
            static[-1] = (static[-2][0],) + static[-1][1:]
assert len(static) == len(dyn), (len(static), len(dyn))
per_line = collections.Counter(); samples = collections.Counter(); alu = collections.Counter()
for (line, f, l, text), d in zip(static, dyn):
    n = int(d["Instructions Executed"] or 0)
    per_line[line] += n
    samples[line] += int(d["# Samples"] or 0)
tot = sum(per_line.values()); ts = sum(samples.values())
print(f"total warp instructions {tot}, samples {ts}")
for line in sorted(k for k in per_line if k is not None):
    if per_line[line] * 1000 >= tot or samples[line] * 1000 >= ts:
        print(f"{top}:{line:4d}  inst {100 * per_line[line] / tot:5.2f}%  samples {100 * samples[line] / ts:5.2f}%")
if len(sys.argv) > 5:
    lo, hi = (int(x) for x in sys.argv[5].split("-"))
    for (line, f, l, text), d in zip(static, dyn):
        if line is not None and lo <= line <= hi:
            print(f"{line:4d} {f[:18]:18s}{l:4d} {int(d['Instructions Executed'] or 0):9d} thr {d['Avg. Threads Executed']:>5s} smp {d['# Samples']:>4s}  {text}")
