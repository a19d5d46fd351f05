#!/usr/bin/env python
"""One-line digest of bench.py logs: python tools/bench_line.py file..."""
import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); c = d.get("config", {}); b = d.get("breakdown", {})
            print(f"{f}: parts={c.get('parts_per_level')} sched={c.get('schedule')}/{c.get('e2e_schedule')} value {d['value']:.0f} ({d['ms_per_step']:.1f} ms)  e2e {d['e2e']['value']:.0f} ({d['e2e'].get('ms_per_step', 0):.1f} ms)  seq {b.get('device_sequential_ms_per_step')}")
            print("    seq:", {k: v["ms"] for k, v in b.get("device_sequential", {}).items()}, " e2e:", {k: v["ms"] for k, v in b.get("e2e", {}).items()})
