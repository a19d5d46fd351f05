#!/usr/bin/env python
"""Aggregate the per-launch csv of tools/ncu_metrics.sh by kernel: python tools/ncu_metrics_summary.py metrics.csv [skip_first_half]
Columns: launches, total ms, warp instructions (G), machine-ms at 100 % issue (592 schedulers x 1.965 GHz), mean issue %, mean resident warps %, DRAM GB."""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; ix = {k: i for i, k in enumerate(h)}
per = collections.OrderedDict()
launches = collections.defaultdict(dict)
for r in rows[1:]:
    if len(r) != len(h): continue
    key = (r[ix["ID"]]); launches[key]["name"] = r[ix["Kernel Name"]]
    try: v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError: continue
    launches[key][r[ix["Metric Name"]]] = v; launches[key]["unit:" + r[ix["Metric Name"]]] = r[ix["Metric Unit"]]
ids = sorted(launches, key=int)
if len(sys.argv) > 2: ids = ids[len(ids) // 2:]   # the second (warm) pass only
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for i in ids:
    L = launches[i]; n = re.sub(r"\(.*", "", L["name"])
    a = agg[n]; a["n"] += 1
    t = L.get("gpu__time_duration.sum", 0.0); u = L.get("unit:gpu__time_duration.sum", "ns")
    t_ms = t / 1e6 if u in ("ns", "nsecond") else (t / 1e3 if u in ("us", "usecond") else t)
    a["ms"] += t_ms; a["inst"] += L.get("smsp__inst_executed.sum", 0.0); a["tinst"] += L.get("smsp__thread_inst_executed.sum", 0.0)
    a["issue_w"] += L.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) * t_ms
    a["warps_w"] += L.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0) * t_ms
    a["dram"] += L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
print(f"{'kernel':44s} {'n':>5s} {'ms':>9s} {'Ginst':>8s} {'ms@100%':>8s} {'lanes':>6s} {'issue%':>7s} {'warps%':>7s} {'DRAM GB':>8s}")
tot = collections.defaultdict(float)
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    floor = a["inst"] / (592 * 1.965e9) * 1e3
    print(f"{n[:44]:44s} {int(a['n']):5d} {a['ms']:9.3f} {a['inst']/1e9:8.3f} {floor:8.3f} {a['tinst']/max(a['inst'],1):6.1f} {a['issue_w']/max(a['ms'],1e-9):7.1f} {a['warps_w']/max(a['ms'],1e-9):7.1f} {a['dram']/1e9:8.2f}")
    tot["ms"] += a["ms"]; tot["floor"] += floor; tot["dram"] += a["dram"]
print(f"{'TOTAL':44s} {'':5s} {tot['ms']:9.3f} {'':8s} {tot['floor']:8.3f} {'':6s} {'':7s} {'':7s} {tot['dram']/1e9:8.2f}")
