#!/usr/bin/env python
"""Per-image device times on the reference's corpus (tests/golden/corpus, full-size PNGs) at levels 1 and 2."""
import os, sys, glob, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from PIL import Image
from oracle import pyoracle as po
from xpng_b200 import Codec
cd = Codec(0)
tot = {}
for p in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "corpus", "*.png"))):
    im = Image.open(p); im = im.convert("RGBA" if im.mode in ("RGBA", "LA", "P") and "transparency" in im.info or im.mode == "RGBA" else "RGB")
    px = po.normalize(np.ascontiguousarray(np.array(im)))
    for lv in (1, 2):
        f = cd.encode(lv, [px])[0]; cd.decode([f])
        cd.encode(lv, [px]); ke = cd.last_kernel_ms
        cd.decode([f]); kd = cd.last_kernel_ms
        mp = px.shape[0] * px.shape[1] / 1e6
        print(f"{os.path.basename(p)[:28]:28s} {px.shape[1]:5d}x{px.shape[0]:<5d}x{px.shape[2]} L{lv} size {len(f):9d}  enc {ke:7.2f} ms ({mp/ke*1e3:8.0f} MPix/s)  dec {kd:7.2f} ms ({mp/kd*1e3:8.0f} MPix/s)")
        t = tot.setdefault(lv, [0, 0, 0]); t[0] += mp; t[1] += ke; t[2] += kd
for lv, (mp, ke, kd) in tot.items():
    print(f"corpus total L{lv}: {mp:.1f} MPix  enc {ke:.1f} ms ({mp/ke*1e3:.0f} MPix/s)  dec {kd:.1f} ms ({mp/kd*1e3:.0f} MPix/s)  [one image per call]")
