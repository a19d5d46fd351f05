#!/usr/bin/env python
"""Golden subset of BASELINE configs[2]: the first 64 sintel-like 1080p frames (seeds 1000..1063) coded by the UNMODIFIED
reference (oracle/_ref, built from /root/reference by oracle/Makefile) at levels 1 and 2.
Writes tests/golden/config3_64.json: per seed and level the .xpng size and sha256."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po          # noqa: E402
from xpng_b200 import synth                # noqa: E402

assert po.ref_available(), "build oracle/_ref first (make -C oracle)"
out = {}
for seed in range(1000, 1064):
    f = synth.sintel_like(seed)
    e = {}
    for lv in (1, 2):
        b = po.ref_encode(lv, f)
        e[str(lv)] = [len(b), hashlib.sha256(b).hexdigest()]
    out[str(seed)] = e
    print(seed, e["1"][0], e["2"][0], flush=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "config3_64.json"), "w"), indent=0, sort_keys=True)
