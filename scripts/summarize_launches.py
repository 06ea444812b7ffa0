"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip_until = sys.argv[2] if len(sys.argv) > 2 else None
rows = []
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith('==')]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
order = []
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    name = re.sub(r'^void ', '', name)[:90]
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    us = v / 1000.0 if unit in ('ns', 'nsecond') else v * (1000.0 if unit in ('ms', 'msecond') else 1.0)
    tot[name][0] += 1
    tot[name][1] += us
total = sum(v[1] for v in tot.values())
print(f'# {path}: {sum(v[0] for v in tot.values())} launches, {total / 1000:.3f} ms total (cold-cache, serialised: compare shares)')
print('| kernel | launches | total us | share |')
print('|---|---:|---:|---:|')
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{k}` | {n} | {us:.1f} | {us / total * 100:.1f} % |')
