#!/usr/bin/env python
"""Summarise raw-page csv exports of ncu --set full captures (tools/ncu_batch_csv.sh): python tools/ncu_csv_summary.py file_raw.csv ...
Also writes profiles/r02_traffic.json-style records on request: --traffic <pixels per launch> (stdout json)."""
import csv, json, re, sys
KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"), ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes_per_inst"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"), ("lts__t_bytes.sum", "l2_bytes"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall_long_sb"), ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall_short_sb"),
        ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall_wait"), ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall_barrier"),
        ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall_math"), ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall_not_selected"),
        ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall_mio"), ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall_lg")]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
args = [a for a in sys.argv[1:] if not a.startswith("--")]
traffic = None
if "--traffic" in sys.argv: traffic = float(sys.argv[sys.argv.index("--traffic") + 1]); args = [a for a in args if a != sys.argv[sys.argv.index("--traffic") + 1]]
recs = {}
for f in args:
    rows = list(csv.reader(open(f)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r)); u = dict(zip(h, units))
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("xpb::", "").replace("(int)", "")
        vals = {}
        for k, short in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                try: vals[short] = float(d[k].replace(",", "")) * SCALE.get(u[k], 1)
                except ValueError: pass
        if traffic is None:
            print(f"== {name}  ({f.split('/')[-1]})")
            print("   " + "  ".join(f"{k}={v:.4g}" for k, v in vals.items()))
            if "dram_read" in vals: print(f"   DRAM {(vals['dram_read'] + vals['dram_write'])/1e9:.3f} GB in {vals['duration']:.3f} ms = {(vals['dram_read'] + vals['dram_write'])/1e6/vals['duration']:.0f} GB/s; "
                                          f"machine time of its instructions at 100 % issue: {vals.get('warp_inst', 0)/(592*1.965e9)*1e3:.3f} ms")
        elif name not in recs and "dram_read" in vals:
            recs[name] = {"dram_bytes_per_launch": int(vals["dram_read"] + vals["dram_write"]), "pixels_per_launch": int(traffic),
                          "dram_bytes_per_pixel": round((vals["dram_read"] + vals["dram_write"]) / traffic, 4), "duration_ms": round(vals["duration"], 3), "report": f.split("/")[-1]}
if traffic is not None: json.dump(recs, sys.stdout, indent=1); print()
