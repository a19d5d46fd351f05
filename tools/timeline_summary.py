#!/usr/bin/env python
"""Summarise the TL lines of tools/batch_timeline.py: per call and kernel the first start, last end, sum of durations,
and per lane the order of events: python tools/timeline_summary.py timeline.txt"""
import sys, collections
level = None; calls = collections.OrderedDict()
for l in open(sys.argv[1]):
    p = l.split()
    if len(p) >= 3 and p[0] == "TL" and p[1] == "level": level = p[2]; continue
    if len(p) == 6 and p[0] == "TL":
        calls.setdefault((level, p[1]), []).append((int(p[2]), p[3], float(p[4]), float(p[5])))
for (lv, what), rows in calls.items():
    end = max(r[3] for r in rows)
    print(f"== L{lv} {what}: {len(rows)} launches, span {end:.2f} ms")
    agg = collections.OrderedDict()
    for lane, name, a, b in rows:
        g = agg.setdefault(name, [1e9, 0, 0, 0]); g[0] = min(g[0], a); g[1] = max(g[1], b); g[2] += b - a; g[3] += 1
    for name, g in sorted(agg.items(), key=lambda kv: -kv[1][2])[:14]:
        print(f"   {name:32s} x{g[3]:<3d} first start {g[0]:8.2f}  last end {g[1]:8.2f}  sum of durations {g[2]:8.2f}")
    for lane in sorted({r[0] for r in rows})[:2] + sorted({r[0] for r in rows})[-1:]:
        print(f"   -- lane {lane}: " + " | ".join(f"{n.replace('k_dec_','').replace('k_','')} {a:.1f}-{b:.1f}" for ln, n, a, b in rows if ln == lane and b - a > 0.3))
