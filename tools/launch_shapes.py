"""Per-(kernel, grid) breakdown of an ncu launch list."""
import csv, re, sys
from collections import defaultdict
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(lines):
    if r['Metric Name'] != 'gpu__time_duration.sum':
        continue
    n = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ivv::', '').replace('ivv::', '')
    if n.startswith('void at::') or n.startswith('void <unnamed>'):
        continue
    key = (n, r['Grid Size'])
    agg[key][0] += 1
    agg[key][1] += float(r['Metric Value'].replace(',', ''))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]:44s} grid={k[1]:16s} n={v[0]:3d} total={v[1]/1e6:7.3f} ms avg={v[1]/v[0]/1e3:8.1f} us")
