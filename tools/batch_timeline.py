#!/usr/bin/env python
"""Timeline of one batch call under its real concurrency (xpngb_profile mode 3: events around every launch, no host waits):
   python tools/batch_timeline.py <frames> <levels> > timeline.txt      lines: TL enc|dec lane kernel start_ms end_ms
A summary per kernel name (first start, last end, busy union) follows on stdout."""
import os, sys, collections, subprocess
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpng_b200
from xpng_b200 import synth, Codec
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 256
levels = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "12")]
imgs = synth.sintel_batch(range(1000, 1000 + nf))
cd = Codec(0); lib = xpng_b200.lib()
shapes = [a.shape for a in imgs]
descs, total = Codec.layout(shapes)
buf = np.zeros(total + 64, np.uint8)
for d, a in zip(descs, imgs): buf[d.offset:d.offset + a.size] = a.reshape(-1)
cap = int(lib.xpngb_encode_bound(descs, nf))
d_px = torch.from_numpy(buf).cuda(); d_f = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda"); d_back = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
def enc(lv):
    d, _ = Codec.layout(shapes); return cd.encode_raw(lv, d, nf, d_px.data_ptr(), total, 1, d_f.data_ptr(), cap, 1)
def dec(offs, sz):
    d, _ = Codec.layout(shapes)
    for x in d: x.w = x.h = 0
    cd.decode_raw(d, nf, d_f.data_ptr(), cap, 1, offs, sz, d_back.data_ptr(), total, 1)
for lv in levels:
    offs, sz = enc(lv); dec(offs, sz)
    offs, sz = enc(lv); e_ms = cd.last_kernel_ms; dec(offs, sz); d_ms = cd.last_kernel_ms
    print(f"# L{lv} {nf} frames: enc {e_ms:.2f} ms dec {d_ms:.2f} ms (plain)", flush=True)
    sys.stderr.write(f"TL level {lv}\n"); sys.stderr.flush()
    cd.profile(3)
    offs, sz = enc(lv); e_ms = cd.last_kernel_ms; dec(offs, sz); d_ms = cd.last_kernel_ms
    cd.profile(0)
    print(f"# L{lv} {nf} frames: enc {e_ms:.2f} ms dec {d_ms:.2f} ms (timeline mode)", flush=True)
