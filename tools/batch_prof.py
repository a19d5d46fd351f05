#!/usr/bin/env python
"""Per-kernel times (serialised) of one batch call: python tools/batch_prof.py [frames] [level]"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpng_b200
from xpng_b200 import synth, Codec
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 256
levels = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "1")]
imgs = synth.sintel_batch(range(1000, 1000 + nf))
cd = Codec(0); lib = xpng_b200.lib()
shapes = [a.shape for a in imgs]
descs, total = Codec.layout(shapes)
buf = np.zeros(total + 64, np.uint8)
for d, a in zip(descs, imgs): buf[d.offset:d.offset + a.size] = a.reshape(-1)
cap = int(lib.xpngb_encode_bound(descs, nf))
d_px = torch.from_numpy(buf).cuda(); d_f = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda"); d_back = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
for lv in levels:
  def enc():
      d, _ = Codec.layout(shapes); return cd.encode_raw(lv, d, nf, d_px.data_ptr(), total, 1, d_f.data_ptr(), cap, 1)
  def dec(offs, sz):
      d, _ = Codec.layout(shapes)
      for x in d: x.w = x.h = 0
      cd.decode_raw(d, nf, d_f.data_ptr(), cap, 1, offs, sz, d_back.data_ptr(), total, 1)
  offs, sz = enc(); dec(offs, sz)
  if len(sys.argv) > 3 and sys.argv[3] == "noprof":   # under ncu: one more plain pass, no event timing
      offs, sz = enc(); dec(offs, sz); continue
  for what in ("enc", "dec"):
      cd.profile(True)
      if what == "enc": offs, sz = enc()
      else: dec(offs, sz)
      rep = cd.profile_report(); cd.profile(False)
      print(f"--- {nf} x 1080p L{lv} {what}: {sum(v[0] for v in rep.values()):.2f} ms serialised")
      for k, (ms, c) in sorted(rep.items(), key=lambda kv: -kv[1][0])[:9]: print(f"    {k:34s} {ms:9.3f} ms x{c}")
