#!/usr/bin/env python3
"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches) of the
kernels in `ncu --set full` captures -> the JSON bench.py reads for its roofline.traffic entries.
usage: make_traffic_json.py out.json source-note rep [rep...]"""
import csv
import json
import subprocess
import sys

out, note, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
UNIT = {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {"source": note}
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    acc = {}
    for r in rows[2:]:
        name = r[ik].split("(")[0].replace("void ", "").split("::")[-1].split("<")[0]
        b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        acc.setdefault(name, []).append(b)
    for k, v in acc.items():
        res[k] = round(sum(v) / len(v))
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
