#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("mb200::", "")
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"# {path}: {sum(cnt.values())} launches, {T / 1e6:.3f} ms of kernel time (cold-cache, serialised: compare shares)")
print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
for k, v in tot.most_common(25):
    print(f"{k[:44]:44s} {cnt[k]:8d} {v / 1e6:10.3f} {v / cnt[k] / 1e3:9.2f} {100 * v / T:6.1f}%")
