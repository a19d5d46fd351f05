#!/usr/bin/env python
"""Batch throughput probe (BASELINE configs 3, 4, 5): device-resident encode/decode of a batch in ONE call.
Usage: python tools/batch_bench.py [frames]   (env XPNGB_LAT_MAX_BLOCKS selects the kernel family)"""
import ctypes as C, sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpng_b200
from xpng_b200 import synth, Codec

def run(name, imgs, levels, reps=3):
    cd = Codec(0); lib = xpng_b200.lib()
    shapes = [a.shape for a in imgs]
    descs, total = Codec.layout(shapes)
    buf = np.zeros(total + 64, np.uint8)
    for d, a in zip(descs, imgs): buf[d.offset:d.offset + a.size] = a.reshape(-1)
    cap = int(lib.xpngb_encode_bound(descs, len(imgs)))
    d_px = torch.from_numpy(buf).cuda(); d_f = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda"); d_back = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    npx = sum(a.shape[0] * a.shape[1] for a in imgs)
    for lv in levels:
        best_e = best_d = 1e9; ke = kd = 0
        for r in range(reps):
            d, _ = Codec.layout(shapes)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            offs, sz = cd.encode_raw(lv, d, len(imgs), d_px.data_ptr(), total, 1, d_f.data_ptr(), cap, 1)
            te = time.perf_counter() - t0; ke = cd.last_kernel_ms
            d2, _ = Codec.layout(shapes)
            for x in d2: x.w = x.h = 0
            torch.cuda.synchronize(); t0 = time.perf_counter()
            cd.decode_raw(d2, len(imgs), d_f.data_ptr(), cap, 1, offs, sz, d_back.data_ptr(), total, 1)
            td = time.perf_counter() - t0; kd = cd.last_kernel_ms
            best_e = min(best_e, te); best_d = min(best_d, td)
        ok = bool(torch.equal(d_back[:total], d_px[:total])) if all(a.shape[2] == 3 for a in imgs) else None
        xb = int(sum(sz)); raw = sum(a.size for a in imgs)
        print(f"{name:28s} L{lv} n={len(imgs):4d} {npx/1e6:8.1f} MPix  enc {npx/1e6/best_e:9.1f} MPix/s ({(raw+xb)/1e9/best_e:7.1f} GB/s, k {ke:7.2f} ms)  "
              f"dec {npx/1e6/best_d:9.1f} MPix/s ({(raw+xb)/1e9/best_d:7.1f} GB/s, k {kd:7.2f} ms)  ratio {raw/xb:5.2f} roundtrip={ok}", flush=True)

if __name__ == "__main__":
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    which = sys.argv[2] if len(sys.argv) > 2 else "345"
    if "3" in which: run(f"cfg3 1080p sintel-like x{nf}", synth.sintel_batch(range(1000, 1000 + nf)), (1, 2))
    if "5" in which: run("cfg5 gray 4096^2 x4", [synth.gray_as_rgb(4096, 4096, 3000 + i) for i in range(4)], (2, 1))
    if "4" in which: run("cfg4 rgba 8192^2", [synth.rgba(8192, 8192, 2)], (1,))
