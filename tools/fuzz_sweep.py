#!/usr/bin/env python
"""Corrupt-input sweep: mutated .xpng files must make xpngb_decode fail cleanly or return garbage, never fault or hang.
After every group of mutations a clean file is decoded; a CUDA fault would poison the context and show up there.
Usage: python tools/fuzz_sweep.py [mutations] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po
from xpng_b200 import synth, Codec

def mutate(rng, good):
    f = bytearray(good); kind = int(rng.integers(0, 6))
    if kind == 0:
        for _ in range(int(rng.integers(1, 8))): f[int(rng.integers(8, len(f)))] = int(rng.integers(0, 256))
    elif kind == 1:
        for _ in range(int(rng.integers(1, 4))): f[int(rng.integers(8, len(f)))] ^= 1 << int(rng.integers(0, 8))
    elif kind == 2: f = f[: int(rng.integers(11, len(f)))]                     # truncation
    elif kind == 3:                                                            # corrupt a 32-bit field near the start of a tile
        p = int(rng.integers(8, min(len(f) - 4, 64))); f[p:p + 4] = rng.integers(0, 256, 4, dtype=np.uint8).tobytes()
    elif kind == 4:                                                            # overwrite a run with one byte
        p = int(rng.integers(8, len(f))); n = int(rng.integers(1, 200)); f[p:p + n] = bytes([int(rng.integers(0, 256))]) * len(f[p:p + n])
    else:                                                                      # splice another region
        p, q, n = int(rng.integers(8, len(f))), int(rng.integers(8, len(f))), int(rng.integers(4, 400))
        f[p:p + n] = f[q:q + n][: len(f[p:p + n])]
    return bytes(f)

if __name__ == "__main__":
    nmut = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    imgs = [synth.rgb(300, 340, 1), synth.rgba(280, 300, 2), synth.gray_as_rgb(200, 260, 3), synth.rgb(500, 520, 4), synth.noise(64, 64, 5)]
    goods = [(lv, im, po.encode(lv, im)) for im in imgs for lv in (1, 2)]
    cd = Codec(0); t0 = time.time(); fails = ok = 0
    for i in range(nmut):
        lv, im, good = goods[int(rng.integers(len(goods)))]
        batch = [mutate(rng, good) for _ in range(int(rng.integers(1, 4)))]
        try:
            cd.decode(batch); ok += 1
        except (RuntimeError, ValueError):
            fails += 1
        if i % 50 == 49:
            lv, im, good = goods[int(rng.integers(len(goods)))]
            assert np.array_equal(cd.decode([good])[0], po.normalize(im)), "context poisoned or wrong decode after corrupt input"
    print(f"fuzz sweep: {nmut} corrupt batches, {fails} rejected, {ok} decoded to garbage, clean decodes stayed correct, {time.time() - t0:.0f} s")
