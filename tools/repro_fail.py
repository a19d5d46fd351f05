import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po
from xpng_b200 import Codec
z = np.load(sys.argv[1]); lv = int(z["level"]); batch = [z[k] for k in z.files if k != "level"]
cd = Codec(0); want = [po.encode(lv, im) for im in batch]
back = cd.decode(want)
print("decode ok", all(np.array_equal(b, po.normalize(im)) for b, im in zip(back, batch)))
print("encode ok", cd.encode(lv, batch) == want)
