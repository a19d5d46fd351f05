#!/usr/bin/env python
"""End-to-end (pinned host buffers, copies inside the timed region) step of the frame batch under different job cuts and
submission orders: python tools/e2e_batch_probe.py [frames] [configs]     configs: comma list of PARTSxMODE, e.g. 1xchained,4xchained,4xfree,4xlag2
  chained: job i encodes when job i-1 has encoded;  free: all jobs start at once;  lagK: job i encodes when job i-K has encoded
A job = one level x one part of the batch: encode, then decode, on its own codec context and host thread (as bench.py)."""
import ctypes as C, os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpng_b200
from xpng_b200 import synth, shard
from concurrent.futures import ThreadPoolExecutor

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
configs = (sys.argv[2] if len(sys.argv) > 2 else "1xchained,2xchained,4xchained,4xlag2,4xfree").split(",")
FW, FH = 1920, 1080
FRAME_B = FW * FH * 3
FILE_B = (8 + FRAME_B + 15) & ~15
lib = xpng_b200.lib()
h_px = torch.empty(F * FRAME_B + 64, dtype=torch.uint8).pin_memory()
hv = h_px.numpy()
for k0 in range(0, F, 64):
    for k, f in enumerate(synth.sintel_batch(range(1000 + k0, 1000 + min(k0 + 64, F)))):
        hv[(k0 + k) * FRAME_B:(k0 + k + 1) * FRAME_B] = f.reshape(-1)
h_files = {lv: torch.empty(F * FILE_B // 2 + 64, dtype=torch.uint8).pin_memory() for lv in (1, 2)}
h_back = {lv: torch.empty(F * FRAME_B + 64, dtype=torch.uint8).pin_memory() for lv in (1, 2)}
shapes = [(FH, FW, 3)] * F
cds = {}

def descs(m, zero=False):
    d = xpng_b200.Codec.layout(shapes[:m])[0]
    if zero:
        for x in d: x.w = x.h = 0
    return d

def run(parts, mode):
    rng = [shard.shard_range(F, p, parts) for p in range(parts)]
    jobs = [(lv, p) for lv in (1, 2) for p in range(parts)]
    for j in list(cds):                                  # contexts keep their scratch: close the ones this cut does not use
        if j not in jobs: cds.pop(j).close()
    for j in jobs:
        if j not in cds: cds[j] = xpng_b200.Codec(0)
    lag = 1 if mode == "chained" else (int(mode[3:]) if mode.startswith("lag") else 0)
    pool = ThreadPoolExecutor(len(jobs))
    def job(i, evs):
        lv, p = jobs[i]; a, b = rng[p]; m = b - a
        if lag and i >= lag: evs[i - lag].wait()
        fbase = a * (FILE_B // 2); fcap = m * (FILE_B // 2)
        try:
            offs, sz = cds[jobs[i]].encode_raw(lv, descs(m), m, h_px.data_ptr() + a * FRAME_B, m * FRAME_B, 0, h_files[lv].data_ptr() + fbase, fcap, 0)
        finally:
            evs[i].set()                                 # a failed job must not leave the others waiting
        cds[jobs[i]].decode_raw(descs(m, True), m, h_files[lv].data_ptr() + fbase, fcap, 0, offs, sz, h_back[lv].data_ptr() + a * FRAME_B, m * FRAME_B, 0)
    def step():
        evs = [threading.Event() for _ in jobs]
        for fu in [pool.submit(job, i, evs) for i in range(len(jobs))]: fu.result()
    step()
    torch.cuda.synchronize(); ts = []
    for _ in range(3):
        t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    ok = all(np.array_equal(h_back[lv].numpy()[:F * FRAME_B], hv[:F * FRAME_B]) for lv in (1, 2))
    print(f"{F} frames, {parts} part(s) per level, {mode:8s}: {np.mean(ts):7.1f} ms per step (min {np.min(ts):7.1f})  {4 * F * FW * FH / 1e6 / (np.mean(ts) / 1e3):8.0f} MPix/s  roundtrip={ok}", flush=True)
    pool.shutdown()

for c in configs:
    p, m = c.split("x")
    run(int(p), m)
