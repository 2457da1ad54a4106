#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: hottest SASS/source lines by warp-stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
si, ci, ei = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = 0; L = []
for r in rows[2:]:
    try: v = int(r[ci])
    except Exception: v = 0
    tot += v
    st = sorted([(int(r[i] or 0), h) for i, h in stall_cols], reverse=True)[:2]
    L.append((v, r[ei], r[si][:95], ",".join(f"{h[6:]}:{n}" for n, h in st if n)))
print("total samples", tot)
for v, e, s, st in sorted(L, reverse=True)[:top]:
    print("%6d %5.1f%% exec=%-9s %-95s %s" % (v, 100 * v / max(tot, 1), e, s, st))
