#!/usr/bin/env python
"""A/B of environment-selected kernel heuristics on one batch, frames generated once:
   python tools/batch_ab.py <frames> "<VAR=val,VAR=val>" "<...>" ...     (an empty string is the default build)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpng_b200
from xpng_b200 import synth, Codec
nf = int(sys.argv[1])
imgs = synth.sintel_batch(range(1000, 1000 + nf))
lib = xpng_b200.lib()
shapes = [a.shape for a in imgs]
descs, total = Codec.layout(shapes)
buf = np.zeros(total + 64, np.uint8)
for d, a in zip(descs, imgs): buf[d.offset:d.offset + a.size] = a.reshape(-1)
cap = int(lib.xpngb_encode_bound(descs, nf))
d_px = torch.from_numpy(buf).cuda(); d_f = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda"); d_back = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
npx = sum(h * w for h, w, _ in shapes) / 1e6
for cfg in sys.argv[2:] or [""]:
    keys = []
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("="); os.environ[k] = v; keys.append(k)
    cd = Codec(0)
    for lv in (1, 2):
        be = bd = 1e9
        for r in range(3):
            d, _ = Codec.layout(shapes)
            offs, sz = cd.encode_raw(lv, d, nf, d_px.data_ptr(), total, 1, d_f.data_ptr(), cap, 1); be = min(be, cd.last_kernel_ms)
            d2, _ = Codec.layout(shapes)
            for x in d2: x.w = x.h = 0
            d_back.zero_()
            cd.decode_raw(d2, nf, d_f.data_ptr(), cap, 1, offs, sz, d_back.data_ptr(), total, 1); bd = min(bd, cd.last_kernel_ms)
        if os.environ.get("BATCH_AB_PROF"):
            cd.profile(True); d, _ = Codec.layout(shapes)
            cd.encode_raw(lv, d, nf, d_px.data_ptr(), total, 1, d_f.data_ptr(), cap, 1)
            rep = cd.profile_report(); cd.profile(False)
            print("    enc serialised:", ", ".join(f"{k} {v[0]:.2f}" for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0])[:6]))
        ok = bool(torch.equal(d_back[:total], d_px[:total]))
        print(f"[{cfg or 'default':40s}] {nf} x 1080p L{lv}: enc {be:8.2f} ms {npx/be*1e3:9.0f} MPix/s   dec {bd:8.2f} ms {npx/bd*1e3:9.0f} MPix/s  roundtrip={ok}", flush=True)
    cd.close()
    for k in keys: del os.environ[k]
