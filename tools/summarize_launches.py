"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total time, share."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, ns))
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"total launches {len(rows)}  total kernel time {tot / 1e6:.3f} ms")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:70]:70s} {c:8d} {ns / 1e6:10.3f} {100 * ns / tot:6.1f}%")
