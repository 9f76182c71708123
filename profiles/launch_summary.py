"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python profiles/launch_summary.py launches.csv [n_steps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
    k = d["Kernel Name"][:100]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot / steps:.1f} us per step over {steps:g} steps")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{a[1] / steps:9.1f} us {a[0] / steps:6.1f} launches {100 * a[1] / tot:5.1f}%  {k}")
