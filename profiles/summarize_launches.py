#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[1:]:
    try:
        d[r[ki][:70]].append(float(r[vi].replace(',', '')))
    except ValueError:
        pass
mine = {k: v for k, v in d.items() if 'psm::' in k}
tot = sum(sum(v) / len(v) * (len(v) / max(len(x) for x in mine.values()) if False else 1) for v in mine.values())
steps = max(len(v) for k, v in mine.items() if 'gather' in k) if any('gather' in k for k in mine) else 1
per_step = {k: sum(v) / steps for k, v in mine.items() if len(v) >= steps // 4}
T = sum(per_step.values())
print('%-72s %5s %10s %10s %7s' % ('kernel', 'n', 'mean_ns', 'ns/step', 'share'))
for k, v in sorted(per_step.items(), key=lambda kv: -kv[1]):
    print('%-72s %5d %10.0f %10.0f %6.1f%%' % (k, len(d[k]), sum(d[k]) / len(d[k]), v, 100 * v / T))
print('sum of own kernels per step: %.1f us over %d steps' % (T / 1e3, steps))
