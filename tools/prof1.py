import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import pyoracle as po
from xpng_b200 import synth, Codec
cd = Codec(0)
img = synth.rgb(2160,3840,1)
f = po.encode(1,img)
for it in range(2):
    print("--- encode", file=sys.stderr); g = cd.encode(1,[img])[0]; assert g==f
    print("--- decode", file=sys.stderr); b = cd.decode([f])[0]; assert np.array_equal(b,img)
img = synth.rgba(2048,2048,2); f = po.encode(1,img)
print("--- rgba encode", file=sys.stderr); g = cd.encode(1,[img])[0]; assert g==f
print("--- rgba decode", file=sys.stderr); b = cd.decode([f])[0]; assert np.array_equal(b,img)
