#!/usr/bin/env python
"""Summarise an ncu --metrics gpu__time_duration.sum launch list (CSV) per kernel."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1e-3)
    name = r[ki].split('(')[0][:56]
    agg[name][0] += 1
    agg[name][1] += v * scale
tot = sum(v[1] for v in agg.values())
print("%-58s %5s %12s %7s %10s" % ("kernel", "n", "total us", "share", "avg us"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-58s %5d %12.1f %6.1f%% %10.2f" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
print("%-58s %5d %12.1f" % ("total", sum(v[0] for v in agg.values()), tot))
