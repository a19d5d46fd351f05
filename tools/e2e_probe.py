#!/usr/bin/env python
"""Where the end-to-end step (host buffers, copies inside the timed region) loses time against the device-resident one:
times the 4K-frame step for several submission orders / level subsets.  python tools/e2e_probe.py"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpng_b200
from xpng_b200 import synth
from concurrent.futures import ThreadPoolExecutor

H, W = 2160, 3840
frame = synth.rgb(H, W, 1)
lib = xpng_b200.lib()
LEVELS = (1, 2, 7)
cds = {lv: xpng_b200.Codec(0) for lv in LEVELS}
descs, total = xpng_b200.Codec.layout([frame.shape])
cap = int(lib.xpngb_encode_bound(descs, 1))
dev = torch.device("cuda", 0)
d_px = torch.cat([torch.from_numpy(frame.reshape(-1)).to(dev), torch.zeros(64, dtype=torch.uint8, device=dev)])
d_files = {lv: torch.zeros(cap + 64, dtype=torch.uint8, device=dev) for lv in LEVELS}
d_back = {lv: torch.zeros(total + 64, dtype=torch.uint8, device=dev) for lv in LEVELS}
h_px = torch.from_numpy(frame.reshape(-1).copy()).pin_memory()
h_files = {lv: torch.zeros(cap + 64, dtype=torch.uint8).pin_memory() for lv in LEVELS}
h_back = {lv: torch.zeros(total + 64, dtype=torch.uint8).pin_memory() for lv in LEVELS}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pool = ThreadPoolExecutor(max_workers=3)
sizes = {}

def pipe(lv, px, files, back, on_dev, delay):
    if delay: time.sleep(delay)
    d = xpng_b200.Codec.layout([frame.shape])[0]
    offs, sz = cds[lv].encode_raw(lv, d, 1, px.data_ptr(), total, on_dev, files[lv].data_ptr(), cap, on_dev)
    d = xpng_b200.Codec.layout([frame.shape])[0]; d[0].w = d[0].h = 0
    off = (C.c_uint64 * 1)(int(offs[0])); s = (C.c_uint64 * 1)(int(sz[0]))
    cds[lv].decode_raw(d, 1, files[lv].data_ptr(), cap, on_dev, off, s, back[lv].data_ptr(), total, on_dev)

def run(name, order, on_dev, delays=None, steps=8):
    px, files, back = (d_px, d_files, d_back) if on_dev else (h_px, h_files, h_back)
    delays = delays or {}
    ts = []
    for i in range(steps + 2):
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter()
        list(pool.map(lambda lv: pipe(lv, px, files, back, on_dev, delays.get(lv, 0)), order))
        ts.append((time.perf_counter() - t0) * 1e3)
    ts = ts[2:]
    print(f"{name:58s} {'dev ' if on_dev else 'host'} mean {np.mean(ts):7.2f} ms  min {np.min(ts):7.2f} ms", flush=True)

if os.environ.get("E2E_QUICK"):
    run("level 2 alone", (2,), 0); run("level 1 alone", (1,), 0); run("level 7 alone", (7,), 0); run("levels 2,1", (2, 1), 0)
    sys.exit(0)
for on_dev in (1, 0):
    run("order 1,2,7", (1, 2, 7), on_dev)
    run("order 2,1,7", (2, 1, 7), on_dev)
    run("level 2 alone", (2,), on_dev)
    run("level 1 alone", (1,), on_dev)
    run("level 7 alone", (7,), on_dev)
    run("levels 2,1", (2, 1), on_dev)
    run("order 2,1,7, L1 +0.7 ms, L7 +1.5 ms", (2, 1, 7), on_dev, {1: 0.0007, 7: 0.0015})
    run("order 2,1,7, L7 +3 ms", (2, 1, 7), on_dev, {7: 0.003})
