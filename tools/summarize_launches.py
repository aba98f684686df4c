"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (runs on the GPU box or here)."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(list)
for r in rows[h + 1:]:
    if len(r) == len(hdr):
        agg[r[ik]].append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print("kernel,launches,mean_ns,total_ns,share")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:90]},{len(v)},{sum(v) / len(v):.0f},{sum(v):.0f},{sum(v) / tot:.4f}")
