#!/usr/bin/env python
"""BASELINE.json configs side by side: this library (one batch call, device-resident) vs the unmodified reference
(oracle/_ref/libxpng_ref.so, all host threads, files on tmpfs).  Prints a markdown table.
Usage: python tools/config_report.py [frames_for_config3]"""
import ctypes as C, glob, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xpng_b200
from xpng_b200 import synth, Codec
from oracle import pyoracle as po
sys.path.insert(0, ROOT)
import bench   # reuse _Xpng / _quiet_call

def ref_times(imgs, level, reps=2):
    so = os.path.join(ROOT, "oracle", "_ref", "libxpng_ref.so")
    if not os.path.exists(so): return None, None
    L = C.CDLL(so)
    L.xpng_store_T.restype = C.c_bool; L.xpng_store_T.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(bench._Xpng), C.c_char_p]
    L.xpng_load_T.restype = C.c_bool; L.xpng_load_T.argtypes = [C.c_uint64, C.c_char_p, C.POINTER(bench._Xpng)]
    libc = C.CDLL(None); libc.free.argtypes = [C.c_void_p]
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    paths = [os.path.join(tmp, f"_cfg_{os.getpid()}_{i}.xpng").encode() for i in range(len(imgs))]
    def enc():
        for a, p in zip(imgs, paths):
            pm = bench._Xpng(a.ctypes.data, a.shape[1], a.shape[0], a.size, a.shape[2] == 4)
            assert not L.xpng_store_T(0, level, C.byref(pm), p)
    def dec():
        for p in paths:
            out = bench._Xpng(); assert not L.xpng_load_T(0, p, C.byref(out)); libc.free(C.c_void_p(out.p))
    te = td = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); bench._quiet_call(enc); te = min(te, time.perf_counter() - t0)
        t0 = time.perf_counter(); bench._quiet_call(dec); td = min(td, time.perf_counter() - t0)
    for p in paths:
        if os.path.exists(p): os.remove(p)
    return te, td

def gpu_times(imgs, level, reps=3):
    cd = Codec(0); lib = xpng_b200.lib()
    shapes = [a.shape for a in imgs]
    descs, total = Codec.layout(shapes)
    buf = np.zeros(total + 64, np.uint8)
    for d, a in zip(descs, imgs): buf[d.offset:d.offset + a.size] = a.reshape(-1)
    cap = int(lib.xpngb_encode_bound(descs, len(imgs)))
    d_px = torch.from_numpy(buf).cuda(); d_f = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda"); d_b = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    te = td = 1e9; xb = 0
    for _ in range(reps):
        d, _ = Codec.layout(shapes)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        offs, sz = cd.encode_raw(level, d, len(imgs), d_px.data_ptr(), total, 1, d_f.data_ptr(), cap, 1); te = min(te, time.perf_counter() - t0)
        d2, _ = Codec.layout(shapes)
        for x in d2: x.w = x.h = 0
        torch.cuda.synchronize(); t0 = time.perf_counter()
        cd.decode_raw(d2, len(imgs), d_f.data_ptr(), cap, 1, offs, sz, d_b.data_ptr(), total, 1); td = min(td, time.perf_counter() - t0)
        xb = int(sum(sz))
    cd.close()
    return te, td, xb

def corpus():
    from PIL import Image
    out = []
    for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "corpus", "*.png"))):
        im = Image.open(p); im = im.convert("RGBA" if (im.mode in ("RGBA", "LA") or "transparency" in im.info) else "RGB")
        out.append(po.normalize(np.ascontiguousarray(np.array(im))))
    return out

if __name__ == "__main__":
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cfgs = [("0: corpus images/*.png (17 files, one call)", corpus(), (1, 2)),
            ("1: one 3840x2160 RGB frame", [synth.rgb(2160, 3840, 1)], (1, 2)),
            (f"2: {nf} x 1080p sintel-like frames (of 1000)", [synth.sintel_like(1000 + i) for i in range(nf)], (1, 2)),
            ("3: one 8192x8192 RGBA image", [synth.rgba(8192, 8192, 2)], (1,)),
            ("4: 4 x 4096x4096 grey-as-RGB", [synth.gray_as_rgb(4096, 4096, 3000 + i) for i in range(4)], (2, 1))]
    print(f"| config | level | MPix | B200 enc MPix/s | B200 dec MPix/s | ref enc MPix/s ({os.cpu_count()} thr) | ref dec MPix/s | enc x | dec x | GPU (raw+xpng) GB/s enc / dec |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for name, imgs, levels in cfgs:
        if not imgs: continue
        mp = sum(a.shape[0] * a.shape[1] for a in imgs) / 1e6; raw = sum(a.size for a in imgs)
        for lv in levels:
            ge, gd, xb = gpu_times(imgs, lv)
            re_, rd = ref_times(imgs, lv)
            r = lambda t: f"{mp / t:.0f}" if t else "-"
            print(f"| {name} | {lv} | {mp:.1f} | {mp / ge:.0f} | {mp / gd:.0f} | {r(re_)} | {r(rd)} | {re_ / ge if re_ else 0:.1f} | {rd / gd if rd else 0:.1f} | "
                  f"{(raw + xb) / 1e9 / ge:.1f} / {(raw + xb) / 1e9 / gd:.1f} |", flush=True)
