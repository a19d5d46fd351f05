#!/usr/bin/env python
"""Per-kernel (serialised) times of one batch call over the reference corpus: python tools/corpus_prof.py [level]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.config_report import corpus
from xpng_b200 import Codec
lv = int(sys.argv[1]) if len(sys.argv) > 1 else 2
imgs = corpus(); cd = Codec(0)
f = cd.encode(lv, imgs); cd.decode(f)
for what in ("enc", "dec"):
    cd.profile(True)
    if what == "enc": cd.encode(lv, imgs)
    else: cd.decode(f)
    rep = cd.profile_report(); cd.profile(False)
    print(f"--- corpus L{lv} {what}: {sum(v[0] for v in rep.values()):.2f} ms serialised")
    for k, (ms, c) in sorted(rep.items(), key=lambda kv: -kv[1][0])[:8]: print(f"    {k:34s} {ms:9.3f} ms x{c}")
