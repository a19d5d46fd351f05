"""Device time of k_orient on one 3840x2160 RGB frame and on 64 1080p frames (device-resident both sides),
against the bytes it must move (one read + one write of the raw pixels)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpng_b200
from xpng_b200 import synth
from xpng_b200.codec import OPS, lib

cd = xpng_b200.Codec(0)
peak = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
for name, shapes in (("1x4K", [(2160, 3840, 3)]), ("64x1080p", [(1080, 1920, 3)] * 64), ("1x8192^2 RGBA", [(8192, 8192, 4)])):
    descs, total = cd.layout(shapes)
    src = torch.randint(0, 256, (total + 16,), dtype=torch.uint8, device="cuda")
    dst = torch.empty_like(src)
    for op in ("mv", "mh", "mvh", "r90", "r270"):
        best = 1e9
        for _ in range(6):
            d2, _t = cd.layout(shapes)
            if lib().xpngb_transform(cd._h, OPS[op], d2, len(shapes), C.c_void_p(src.data_ptr()), total, 1, C.c_void_p(dst.data_ptr()), 1):
                raise RuntimeError(cd._err())
            best = min(best, cd.last_kernel_ms)
        print(f"{name:14s} {op:5s} {best*1e3:8.1f} us  {2*total/best/1e6:8.1f} GB/s")
