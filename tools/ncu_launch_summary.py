#!/usr/bin/env python3
"""Per-kernel shares of ONE pass of the path from an `ncu --metrics gpu__time_duration.sum --csv` launch list of
bench.py: the last pass = the launches from the last k_join pair up to the last k_em* launch."""
import csv, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ni, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
L = []
for r in rows:
    name = re.sub(r"\(.*", "", r[ni]).replace("colate::", "")
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r[ui]]
    L.append((name, v))
em = max(i for i, (n, _) in enumerate(L) if n.startswith("k_em"))
bounds = [i for i, (n, _) in enumerate(L[:em]) if n == "k_join_bounds"]
joins = [i for i, (n, _) in enumerate(L[:em]) if n == "k_join"]
if len(bounds) >= 2 and bounds[-1] - bounds[-2] <= 2:       # k_join_bounds, k_join once per genome
    start = bounds[-2]
elif bounds:
    start = bounds[-1]
else:
    start = joins[-2] if len(joins) >= 2 and joins[-1] - joins[-2] <= 2 else joins[-1]
one = L[start:em + 1]
agg = OrderedDict()
for n, v in one:
    c, t = agg.get(n, (0, 0.0))
    agg[n] = (c + 1, t + v)
tot = sum(t for _, t in agg.values())
print("| kernel | launches | us | share |\n|---|---|---|---|")
for n, (c, t) in agg.items():
    print("| %s | %d | %.1f | %.1f %% |" % (n, c, t, 100 * t / tot))
print("| total | %d | %.1f | |" % (len(one), tot))
