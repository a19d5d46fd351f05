#!/usr/bin/env python
"""Per-SASS-instruction stall view of an .ncu-rep (--import-source on): python tools/ncu_sass.py file.ncu-rep [top]
   -> the instructions with the most stall samples, their dominant stall reasons and executed counts"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(out)))
hdr = None; ins = []
for r in rows:
    if len(r) > 10 and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        def f(k):
            try: return float(d.get(k, "0") or 0)
            except ValueError: return 0.0
        st = sorted(((f(k), k[6:]) for k in hdr if k.startswith("stall_")), reverse=True)[:2]
        ins.append((f("# Samples"), f("Instructions Executed"), d["Address"], d["Source"][:70], st))
tot = sum(i[0] for i in ins) or 1
print(f"total samples {tot:.0f}")
for i in sorted(ins, key=lambda i: -i[0])[:top]:
    print(f"{100*i[0]/tot:6.2f}%  x{i[1]/1e6:8.2f}M  {i[2][-5:]}  {i[3]:70s} {i[4][0][1]}={i[4][0][0]:.0f} {i[4][1][1]}={i[4][1][0]:.0f}")
