#!/usr/bin/env python3
"""Hottest SASS lines (warp-stall samples) of one kernel from `ncu -i rep --page source --csv` output.
usage: ncu_src.py source.csv <kernel substring> [top]   (the csv may hold several kernels: sections start with a "Kernel Name" row)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(starts, starts[1:]):
    name = rows[a][1]
    if want not in name: continue
    hdr = rows[a + 1]
    si, ci, ei = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = 0; L = []; agg = {}
    for r in rows[a + 2:b]:
        if len(r) < len(hdr): continue
        try: v = int(r[ci])
        except ValueError: v = 0
        tot += v
        st = []
        for i, h in stall_cols:
            try: n = int(r[i] or 0)
            except ValueError: n = 0
            agg[h] = agg.get(h, 0) + n
            if n: st.append((n, h))
        st.sort(reverse=True)
        L.append((v, r[ei], r[si][:95], ",".join(f"{h[6:]}:{n}" for n, h in st[:2])))
    print("kernel", name[:100]); print("total samples", tot, "SASS lines", len(L))
    print("stall mix:", ", ".join(f"{h[6:]} {100*n/max(1,sum(agg.values())):.1f}%" for h, n in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for v, e, s, st in sorted(L, reverse=True)[:top]:
        print("%6d %5.1f%% exec=%-9s %-95s %s" % (v, 100 * v / max(tot, 1), e, s, st))
    break
