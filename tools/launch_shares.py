#!/usr/bin/env python
"""Per-kernel share of the summed kernel time from an ncu launch list (--metrics gpu__time_duration.sum --csv).
Usage: python tools/launch_shares.py launches.csv [skip_first_n_launches]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in rows if int(r[0]) >= skip and "xpb::" in r[4] or "k_" in r[4]]
rows = [r for r in rows if "at::" not in r[4]]
tot = collections.defaultdict(lambda: [0.0, 0])
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("xpb::", "")
    tot[name][0] += float(r[-1]) / 1e6; tot[name][1] += 1
s = sum(v[0] for v in tot.values())
print(f"| kernel | launches | total ms | ms / launch | share of summed kernel time |\n|---|---|---|---|---|")
for k, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"| `{k}` | {n} | {ms:.3f} | {ms / n:.4f} | {100 * ms / s:.1f} % |")
print(f"\nsum {s:.2f} ms over {sum(v[1] for v in tot.values())} launches")
