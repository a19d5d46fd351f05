#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a few lines: duration, DRAM bytes, issue utilisation, top stall reasons,
and the hottest source lines.  Usage: python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    def g(k, default="?"):
        return d.get(k, default)
    print(f"kernel: {g('Kernel Name')}  grid {g('launch__grid_size')} x block {g('launch__block_size')}, regs {g('launch__registers_per_thread')}, "
          f"dyn smem {g('launch__shared_mem_per_block_dynamic')} B")
    print(f"  duration {g('gpu__time_duration.sum')} {u.get('gpu__time_duration.sum')}   SM clock {g('sm__cycles_elapsed.avg.per_second')} {u.get('sm__cycles_elapsed.avg.per_second')}")
    print(f"  dram read {g('dram__bytes_read.sum')} {u.get('dram__bytes_read.sum')}, write {g('dram__bytes_write.sum')} {u.get('dram__bytes_write.sum')}")
    print(f"  issue slots busy {g('smsp__issue_active.avg.pct_of_peak_sustained_active')} %   inst/cycle/SM {g('sm__inst_executed.avg.per_cycle_active')}   "
          f"warps active {g('sm__warps_active.avg.pct_of_peak_sustained_active')} % of peak")
    stalls = sorted(((float(v or 0), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")), reverse=True)
    print("  stall cycles per issued instruction: " + ", ".join(f"{k.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, k in stalls[:6]))
