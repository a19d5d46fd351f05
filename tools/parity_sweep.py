#!/usr/bin/env python
"""Randomised parity sweep on the GPU box: many shapes / contents / levels, batches of mixed images, every file
byte-compared with the CPU oracle and every decode pixel-compared.  Usage: python tools/parity_sweep.py [cases] [seed]"""
import glob, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po
from xpng_b200 import synth, Codec

def make(rng, corpus):
    kind = int(rng.integers(0, 10))
    h, w = int(rng.integers(1, 900)), int(rng.integers(1, 1100))
    if kind == 0: return synth.rgb(h, w, int(rng.integers(1 << 30)))
    if kind == 1: return synth.rgba(max(h, 4), max(w, 4), int(rng.integers(1 << 30)))
    if kind == 2: return synth.gray_as_rgb(h, w, int(rng.integers(1 << 30)))
    if kind == 3: return synth.noise(h, w, int(rng.integers(1 << 30)))
    if kind == 4:   # flat regions + gradients: long zero runs, few distinct symbols
        y, x = np.mgrid[0:h, 0:w]
        a = np.stack([(x // 7) % 256, (y // 5) % 256, ((x + y) // 11) % 256], -1).astype(np.uint8)
        a[: h // 2, : w // 3] = rng.integers(0, 256, 3)
        return np.ascontiguousarray(a)
    if kind == 5:   # thin shapes
        return synth.rgb(int(rng.integers(1, 4)), int(rng.integers(1, 3000)), 5) if rng.integers(2) else synth.rgb(int(rng.integers(1, 3000)), int(rng.integers(1, 4)), 6)
    if kind == 6 and corpus:   # random crop of a corpus image
        im = corpus[int(rng.integers(len(corpus)))]
        ch, cw = min(im.shape[0], h + 8), min(im.shape[1], w + 8)
        y0, x0 = int(rng.integers(0, im.shape[0] - ch + 1)), int(rng.integers(0, im.shape[1] - cw + 1))
        c = np.ascontiguousarray(im[y0:y0 + ch, x0:x0 + cw])
        if c.shape[2] == 4 and (c.shape[0] < 4 or c.shape[1] < 4): return synth.rgb(h, w, 9)
        return c
    if kind == 7:   # RGBA with random alpha classes incl. fully transparent / opaque tiles
        a = synth.rgba(max(h, 4), max(w, 4), int(rng.integers(1 << 30)))
        m = int(rng.integers(0, 4))
        if m == 0: a[..., 3] = 255
        if m == 1: a[..., 3] = np.where(rng.integers(0, 2, a.shape[:2]) > 0, 255, 0)
        if m == 2: a[: a.shape[0] // 2] = 0
        return a
    if kind == 8:   # two-colour / low-entropy images (raw and run blocks)
        a = np.zeros((h, w, 3), np.uint8) + rng.integers(0, 256, 3).astype(np.uint8)
        a[rng.integers(0, h, max(1, h // 3)), :] = rng.integers(0, 256, 3)
        return a
    return synth.rgb(min(h, 700), min(w + 300, 1300), int(rng.integers(1 << 30)))

if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
    corpus = []
    try:
        from PIL import Image
        for p in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "corpus", "*.png")))[:8]:
            im = Image.open(p); im = im.convert("RGBA" if (im.mode in ("RGBA", "LA") or "transparency" in im.info) else "RGB")
            corpus.append(po.normalize(np.ascontiguousarray(np.array(im))))
    except Exception as e:
        print("no corpus:", e)
    cd = Codec(0); bad = 0; t0 = time.time(); done = 0
    while done < ncases:
        batch = [make(rng, corpus) for _ in range(int(rng.integers(1, 7)))]
        for lv in (1, 2, 7):
            want = [po.encode(lv, im) for im in batch]
            got = cd.encode(lv, batch)
            if got != want:
                bad += 1; print("ENCODE MISMATCH level", lv, [im.shape for im in batch], [g == w for g, w in zip(got, want)], flush=True)
            try:
                back = cd.decode(want)
            except RuntimeError as e:
                print("DECODE EXCEPTION level", lv, [im.shape for im in batch], str(e)[:200], flush=True)
                np.savez(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "sweep_fail.npz"), *batch, level=lv)
                sys.exit(2)
            for b, im in zip(back, batch):
                n = po.normalize(im)
                if b.shape != n.shape or not np.array_equal(b, n):
                    bad += 1; print("DECODE MISMATCH level", lv, im.shape, flush=True)
        done += len(batch)
    print(f"parity sweep: {done} images x 3 levels, {bad} mismatches, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)
