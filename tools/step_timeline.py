#!/usr/bin/env python
"""Wall-clock timeline of one bench step (bench.py's PIPELINES: level 2 | level 1 then 7), per-call start/end (ms)
next to the device span of the call's kernels (CUDA events inside the library)."""
import ctypes as C, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpng_b200
from xpng_b200 import synth, Codec
from concurrent.futures import ThreadPoolExecutor
LEVELS = (1, 2, 7)
frame = synth.rgb(2160, 3840, 1)
cds = {lv: Codec(0) for lv in LEVELS}
lib = xpng_b200.lib()
descs, total = Codec.layout([frame.shape]); cap = int(lib.xpngb_encode_bound(descs, 1))
d_px = torch.cat([torch.from_numpy(frame.reshape(-1)).cuda(), torch.zeros(64, dtype=torch.uint8, device="cuda")])
d_files = {lv: torch.zeros(cap + 64, dtype=torch.uint8, device="cuda") for lv in LEVELS}
d_back = {lv: torch.zeros(total + 64, dtype=torch.uint8, device="cuda") for lv in LEVELS}
sizes = {}; log = []
PIPELINES = ((2,),) if os.environ.get("ONLY2") else ((2,), (1, 7))
pool = ThreadPoolExecutor(len(PIPELINES))
def enc(lv):
    t0 = time.perf_counter(); d = Codec.layout([frame.shape])[0]
    offs, sz = cds[lv].encode_raw(lv, d, 1, d_px.data_ptr(), total, 1, d_files[lv].data_ptr(), cap, 1)
    sizes[lv] = (int(offs[0]), int(sz[0])); log.append(("enc", lv, t0, time.perf_counter(), cds[lv].last_kernel_ms))
def dec(lv):
    t0 = time.perf_counter(); d = Codec.layout([frame.shape])[0]; d[0].w = d[0].h = 0
    off = (C.c_uint64 * 1)(sizes[lv][0]); sz = (C.c_uint64 * 1)(sizes[lv][1])
    cds[lv].decode_raw(d, 1, d_files[lv].data_ptr(), cap, 1, off, sz, d_back[lv].data_ptr(), total, 1)
    log.append(("dec", lv, t0, time.perf_counter(), cds[lv].last_kernel_ms))
for it in range(4):
    log.clear(); torch.cuda.synchronize(); T0 = time.perf_counter()
    list(pool.map(lambda lvs: [(enc(lv), dec(lv)) for lv in lvs], PIPELINES)); T2 = time.perf_counter()
    print(f"iter {it}: step {1e3*(T2-T0):.2f} ms")
    for what, lv, a, b, k in sorted(log, key=lambda r: r[2]):
        print(f"    {what} L{lv}: start {1e3*(a-T0):7.2f} end {1e3*(b-T0):7.2f} (call {1e3*(b-a):6.2f} ms, device span {k:6.2f} ms)")
