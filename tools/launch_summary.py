#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
i = [k for k, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[i]; ix = {h: k for k, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[i + 1:]:
    if len(r) < len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum': continue
    name = r[ix['Kernel Name']][:78]; v = float(r[ix['Metric Value']].replace(',', ''))
    unit = r[ix['Metric Unit']]
    v *= {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'second': 1e3}.get(unit, 1e-6)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"{k:80s} n={v[0]:4d} total={v[1]:9.2f} ms avg={v[1]/v[0]:8.3f} ms {100*v[1]/tot:5.1f}%")
