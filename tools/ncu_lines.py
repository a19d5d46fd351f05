#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep captured with --import-source on (-lineinfo build):
   python tools/ncu_lines.py file.ncu-rep [top]     -> lines by stall samples, with executed warp instructions and shared-memory wavefronts"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(out)))
hdr = None; fname = ""; lines = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): fname = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = dict(zip(hdr, r))
        def f(k):
            try: return float(d.get(k, "0") or 0)
            except ValueError: return 0.0
        lines.append((f("# Samples"), f("Instructions Executed"), f("L1 Wavefronts Shared"), f("L1 Wavefronts Shared Ideal"), f("L2 Theoretical Sectors Global"), fname, r[0], r[1][:110]))
tot_s = sum(l[0] for l in lines) or 1; tot_i = sum(l[1] for l in lines) or 1
print(f"total samples {tot_s:.0f}, warp instructions {tot_i/1e6:.1f} M")
print(f"{'samples%':>8s} {'inst%':>6s} {'smem wf M':>9s} {'ideal':>7s} {'gsect M':>8s}  file:line  source")
for l in sorted(lines, key=lambda l: -l[0])[:top]:
    print(f"{100*l[0]/tot_s:8.1f} {100*l[1]/tot_i:6.1f} {l[2]/1e6:9.1f} {l[3]/1e6:7.1f} {l[4]/1e6:8.1f}  {l[5]}:{l[6]}  {l[7]}")
