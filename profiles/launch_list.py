#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv --log-file x.csv` launch list.
Usage: python profiles/launch_list.py profiles/r3_launches.csv"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name = r[4].split("(")[0]
    tot[name] += float(r[14]) / 1e3
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':58s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k[:58]:58s} {cnt[k]:8d} {tot[k]:12.1f} {tot[k] / cnt[k]:10.1f} {100 * tot[k] / total:6.1f}%")
