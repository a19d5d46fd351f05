#!/usr/bin/env python
"""Device-resident batch encode / decode times under several settings of the library's environment knobs, frames generated
once: python tools/knob_sweep.py <frames> <levels> "K1=V1,K2=V2" "K1=V3" ...   ("" = defaults).  A context reads its knobs
when it is created, so every setting gets a fresh context (closed before the next one)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpng_b200
from xpng_b200 import synth, Codec
nf = int(sys.argv[1]); levels = [int(c) for c in sys.argv[2]]; settings = sys.argv[3:] or [""]
imgs = synth.sintel_batch(range(1000, 1000 + nf))
lib = xpng_b200.lib()
shapes = [a.shape for a in imgs]
descs, total = Codec.layout(shapes)
buf = np.zeros(total + 64, np.uint8)
for d, a in zip(descs, imgs): buf[d.offset:d.offset + a.size] = a.reshape(-1)
cap = int(lib.xpngb_encode_bound(descs, nf))
d_px = torch.from_numpy(buf).cuda(); d_f = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda"); d_back = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
npx = sum(s[0] * s[1] for s in shapes)
for st in settings:
    kv = dict(x.split("=") for x in st.split(",") if x)
    for k, v in kv.items(): os.environ[k] = v
    cd = Codec(0)
    for lv in levels:
        be = bd = 1e9
        for r in range(3):
            d, _ = Codec.layout(shapes)
            offs, sz = cd.encode_raw(lv, d, nf, d_px.data_ptr(), total, 1, d_f.data_ptr(), cap, 1); be = min(be, cd.last_kernel_ms)
            d2, _ = Codec.layout(shapes)
            for x in d2: x.w = x.h = 0
            d_back.zero_()
            cd.decode_raw(d2, nf, d_f.data_ptr(), cap, 1, offs, sz, d_back.data_ptr(), total, 1); bd = min(bd, cd.last_kernel_ms)
        ok = bool(torch.equal(d_back[:total], d_px[:total]))
        print(f"[{st or 'default':44s}] {nf} x 1080p L{lv}: enc {be:8.2f} ms {npx/1e3/be:9.0f} MPix/s   dec {bd:8.2f} ms {npx/1e3/bd:9.0f} MPix/s  roundtrip={ok}", flush=True)
    cd.close()
    for k in kv: del os.environ[k]
