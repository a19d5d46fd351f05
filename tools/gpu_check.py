#!/usr/bin/env python
"""Developer check on a GPU box: GPU encode vs oracle bytes, GPU decode vs pixels, with tile-level
diagnostics on mismatch.  Usage: python tools/gpu_check.py [levels] [case filter]"""
import os, sys, time, struct
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po
from xpng_b200 import synth, Codec


def tiles_of(data, w, h, pxsz):
    grid = po.tile_grid(w, h, pxsz)
    off, out = 8, []
    for t in grid:
        if off + 4 > len(data):
            break
        sz = struct.unpack_from("<I", data, off)[0] & 0xFFFFFF
        out.append((off, sz, data[off + 3], tuple(int(v) for v in t)))
        off += max(sz, 4)
    return out


def describe_m1(blob, pxsz):
    bsz = struct.unpack_from("<I", blob, 4)[0]
    secs = [("hdr", 0, 4), ("bsz", 4, 8), ("kbits", 8, 4 + bsz)]
    off = 4 + bsz
    for c in range(10 if pxsz == 4 else 9):
        if off + 4 > len(blob): break
        w0 = struct.unpack_from("<I", blob, off)[0]
        t, csz = w0 >> 24, (w0 & 0xFFFFFF) if (w0 >> 24) else 4
        n = struct.unpack_from("<I", blob, off + 4)[0] & 0xFFFFFF if t else 0
        secs.append((f"blk{c}(type{t},n={n},csz={csz})", off, off + csz))
        off += csz
    return secs


def diagnose(a, b, w, h, pxsz, mode):
    print(f"   sizes gpu={len(a)} oracle={len(b)} hdr gpu={a[:8].hex()} oracle={b[:8].hex()}")
    if a[:8] != b[:8] or b[3] == 7 or a[3] == 7:
        return
    ta, tb = tiles_of(a, w, h, pxsz), tiles_of(b, w, h, pxsz)
    for i, (x, y) in enumerate(zip(ta, tb)):
        ba, bb = a[x[0]:x[0] + x[1]], b[y[0]:y[0] + y[1]]
        if ba != bb:
            print(f"   tile {i} {y[3]} differs: gpu(off={x[0]},size={x[1]},m={x[2]:#x}) oracle(off={y[0]},size={y[1]},m={y[2]:#x})")
            if mode == 1 and y[2] and x[2]:
                sa, sb = describe_m1(ba, pxsz), describe_m1(bb, pxsz)
                print("     gpu   :", sa); print("     oracle:", sb)
                k = next((j for j in range(min(len(ba), len(bb))) if ba[j] != bb[j]), None)
                sec = next((s for s in sb if s[1] <= k < s[2]), None) if k is not None else None
                print(f"     first differing byte {k} in {sec}: gpu={ba[k:k+8].hex()} oracle={bb[k:k+8].hex()}")
            return
    print("   all common tiles equal; tile count gpu/oracle", len(ta), len(tb))


def main():
    levels = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "17")]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    cases = [
        ("rgb_64x48", synth.rgb(48, 64, 1)), ("rgb_600x500", synth.rgb(500, 600, 1)), ("rgb_1334x750", synth.rgb(750, 1334, 11)),
        ("rgba_700x520", synth.rgba(520, 700, 2)), ("gray_512", synth.gray_as_rgb(512, 512, 3)), ("noise_500", synth.noise(500, 500, 4)),
        ("rgb_37x1500", synth.rgb(1500, 37, 5)), ("rgb_1500x37", synth.rgb(37, 1500, 6)), ("rgb_3x3", synth.rgb(3, 3, 9)),
        ("rgb_2x1", synth.rgb(1, 2, 8)), ("rgb_1x1", synth.rgb(1, 1, 7)), ("flat", np.full((100, 100, 3), 77, np.uint8)),
        ("flat_rgba", np.full((100, 100, 4), 77, np.uint8)), ("sintel", synth.sintel_like(1000)), ("rgb_4k", synth.rgb(2160, 3840, 1)),
        ("rgba_1500", synth.rgba(1500, 1500, 5)), ("halfflat", None),
    ]
    hf = synth.rgb(900, 900, 23); hf[:, :450] = [10, 200, 30]
    cases[-1] = ("halfflat", hf)
    dirty = synth.rgba(200, 300, 22); dirty[dirty[..., 3] == 0] = [9, 8, 7, 0]
    cases.append(("dirty_rgba", dirty))
    cases.append(("opaque_rgba", np.concatenate([synth.rgb(90, 120, 21), np.full((90, 120, 1), 255, np.uint8)], axis=2)))
    cd = Codec(0)
    bad = 0
    for name, img in cases:
        if flt and flt not in name:
            continue
        for lv in levels:
            want = po.encode(lv, img)
            try:
                t = time.time(); got = cd.encode(lv, [img])[0]; te = time.time() - t
            except Exception as e:
                print(f"{name} L{lv}: ENCODE EXC {e}"); bad += 1; continue
            ok = got == want
            norm = po.normalize(img)
            try:
                t = time.time(); back = cd.decode([want])[0]; td = time.time() - t
                dok = back.shape == norm.shape and np.array_equal(back, norm)
            except Exception as e:
                dok = False; td = 0; print(f"   DECODE EXC {e}")
            print(f"{name:14s} L{lv} enc={'OK ' if ok else 'BAD'} dec={'OK ' if dok else 'BAD'} size={len(want)} enc_ms={te*1e3:.1f}(k {cd.last_kernel_ms:.2f}) dec_ms={td*1e3:.1f}", flush=True)
            if not ok:
                bad += 1; diagnose(got, want, img.shape[1], img.shape[0], norm.shape[2], want[3])
            if not dok:
                bad += 1
                if 'back' in dir() and back.shape == norm.shape:
                    d = np.argwhere((back != norm).any(axis=2))
                    print(f"   decode mismatches: {len(d)} px, first at (y,x)={d[0].tolist() if len(d) else None}, got={back[tuple(d[0])] if len(d) else ''} want={norm[tuple(d[0])] if len(d) else ''}")
    # batch: several images in one call
    if not flt:
        imgs = [synth.rgb(300 + 37 * i, 500 + 91 * i, 40 + i) for i in range(5)] + [synth.rgba(400, 333, 50)]
        for lv in levels:
            got = cd.encode(lv, imgs); want = [po.encode(lv, im) for im in imgs]
            ok = got == want
            back = cd.decode(want); dok = all(np.array_equal(b, po.normalize(im)) for b, im in zip(back, imgs))
            print(f"batch6 L{lv} enc={'OK' if ok else 'BAD'} dec={'OK' if dok else 'BAD'}"); bad += (not ok) + (not dok)
    print("FAILURES:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
