#!/usr/bin/env python
"""Extracts per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel in an .ncu-rep
(one `ncu --set full` capture) into a small JSON file that bench.py reads for `roofline.traffic`."""
import csv
import json
import subprocess
import sys

rep, out, workload, entities = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
kernels = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
    wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
    kernels.append({"kernel": d["Kernel Name"].split("(")[0], "grid": d["Grid Size"], "dram_read_bytes": int(rd),
                    "dram_write_bytes": int(wr), "traffic_bytes": int(rd + wr),
                    "gpu_time_us_under_ncu": float(d["gpu__time_duration.sum"])})
json.dump({"source": rep.split("/")[-1], "capture": "ncu --set full --clock-control none, one launch per row",
           "workload": workload, "entities": entities, "kernels": kernels}, open(out, "w"), indent=1)
print(out, len(kernels), "kernels")
