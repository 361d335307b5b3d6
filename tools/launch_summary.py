#!/usr/bin/env python3
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean, share."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict(); tot = 0.0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum': continue
    k = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('nsb::', '').replace('<unnamed>::', '')
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
    a = agg.setdefault(k + ' grid=' + row['Grid Size'], [0, 0.0, 1e9, 0]); a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v); tot += v
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
print(f"total {tot / steps:.1f} us per step over {steps} step(s)")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:78]:78s} n/step={a[0] / steps:6.1f} us/step={a[1] / steps:8.1f} avg={a[1] / a[0]:7.2f} min={a[2]:7.2f} max={a[3]:7.2f} share={a[1] / tot:.3f}")
