#!/usr/bin/env python
"""Aggregate the per-instruction warp-stall samples of an `ncu --page source --print-source sass --csv` export.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass --kernel-id ::regex:NAME:K > f.csv ; ncu_stalls.py f.csv [ntop]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter(); byop = collections.defaultdict(collections.Counter); insts = collections.Counter(); top = []
def I(x):
    try: return int(x)
    except ValueError: return 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address": continue
    src = r[ix['Source']].strip()
    t = src.split()
    op = (t[1] if src.startswith('@') else t[0]).split('.')[0]
    insts[op] += I(r[ix['Instructions Executed']])
    d = {}
    for s in stalls:
        v = I(r[ix[s]])
        tot[s] += v; byop[op][s] += v
        if v: d[s.replace('stall_', '')] = v
    top.append((I(r[ix['# Samples']]), r[ix['Address']][-5:], src, d))
T = sum(tot.values())
print(rows[0][:2]); print("total samples", T)
for s, v in tot.most_common(): print(f"{s:25s} {v:8d} {100*v/T:5.1f}%")
print("by opcode (samples):")
for op, c in sorted(byop.items(), key=lambda kv: -sum(kv[1].values()))[:16]:
    print(f"{op:8s} inst={insts[op]:12d} samples={sum(c.values()):7d}  ", {k.replace('stall_', ''): v for k, v in c.most_common(4)})
print("total warp-inst", sum(insts.values()), "fp64", insts['DFMA'] + insts['DMUL'] + insts['DADD'], "MUFU", insts['MUFU'])
top.sort(key=lambda t: -t[0])
for t in top[:ntop]: print(t)
