"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
   python tools/launches_summary.py gpurun_out/launches.csv > profiles/rNx_launches_summary.csv"""
import collections
import csv
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot, cnt = collections.Counter(), collections.Counter()
for r in csv.DictReader(rows):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1)
    k = r["Kernel Name"][:90]
    tot[k] += v; cnt[k] += 1
s = sum(tot.values())
print("kernel,launches,total_ns,share")
for k, v in tot.most_common():
    print(f'"{k}",{cnt[k]},{int(v)},{v / s:.4f}')
