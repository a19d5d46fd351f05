import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import pyoracle as po
from xpng_b200 import synth, Codec
cd = Codec(0)
which = sys.argv[1] if len(sys.argv) > 1 else "4k"
img = {"4k": lambda: synth.rgb(2160,3840,1), "gray": lambda: synth.gray_as_rgb(4096,4096,3000), "rgba": lambda: synth.rgba(2048,2048,2)}[which]()
for lv in (1,2):
    f = po.encode(lv,img)
    cd.encode(lv,[img]); cd.decode([f])   # warm
    cd.profile(True); g = cd.encode(lv,[img])[0]; rep = cd.profile_report(); assert g==f
    print(f"--- {which} L{lv} encode: total {sum(v[0] for v in rep.values()):.3f} ms")
    for k,(ms,c) in rep.items(): print(f"    {k:34s} {ms:9.3f} ms x{c}")
    cd.profile(True); b = cd.decode([f])[0]; rep = cd.profile_report(); assert np.array_equal(b,po.normalize(img))
    print(f"--- {which} L{lv} decode: total {sum(v[0] for v in rep.values()):.3f} ms")
    for k,(ms,c) in rep.items(): print(f"    {k:34s} {ms:9.3f} ms x{c}")
    cd.profile(False)
