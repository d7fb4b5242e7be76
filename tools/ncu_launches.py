#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the last full step.
usage: python tools/ncu_launches.py gpurun_out/launches.csv [marker-substring-of-first-kernel-of-a-step]"""
import collections
import csv
import sys

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "gemm_prep"
lines = [l for l in open(path) if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    t = float(row["Metric Value"].replace(",", ""))
    t = t / 1e3 if row["Metric Unit"] == "ns" else (t * 1e3 if row["Metric Unit"] == "ms" else t)
    seq.append((row["Kernel Name"][:70], t))
idx = [i for i, (n, _t) in enumerate(seq) if marker in n]
print(len(seq), "launches;", len(idx), "steps")
last = seq[idx[-2]:idx[-1]] if len(idx) >= 2 else seq
tot = sum(t for _n, t in last)
for n, t in last:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  {n}")
print(f"step total {tot:.1f} us")
agg = collections.Counter()
for n, t in last:
    agg[n.split('(')[0]] += t
print("--- by kernel")
for n, t in agg.most_common():
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  {n}")
