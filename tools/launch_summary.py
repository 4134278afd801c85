#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the last step."""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
body = rows[1:]
enc = [n for n, r in enumerate(body) if r[ik].startswith('k_encode')]
sel = body[enc[-1]:] if enc else body
agg = collections.OrderedDict()
for r in sel:
    k = r[ik].split('(')[0]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[iv])
tot = sum(v[1] for v in agg.values())
print("%d launches in the last step, %.3f ms" % (len(sel), tot / 1e6))
for k, v in agg.items():
    print("%-32s n=%4d %10.3f ms %5.1f%%" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))
if len(sys.argv) > 2:
    for r in sel:
        if sys.argv[2] in r[ik]: print(r[ik][:28], "%.3f" % (float(r[iv]) / 1e6))
