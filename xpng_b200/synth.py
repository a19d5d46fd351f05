"""Seeded synthetic inputs of the BASELINE.json shapes; integer-only numpy.

Every generator returns a C-contiguous uint8 array of shape (h, w, 3) or (h, w, 4).
"""
import numpy as np


def _smooth(rng, h, w, c):
    """Coarse random grid, bilinear x64 upsample in integer arithmetic."""
    g = rng.integers(0, 256, (h // 64 + 2, w // 64 + 2, c)).astype(np.int32)
    ys, xs = np.arange(h), np.arange(w)
    gy, fy = (ys // 64)[:, None, None], (ys % 64)[:, None, None]
    gx, fx = (xs // 64)[None, :, None], (xs % 64)[None, :, None]
    gy, gx = gy[:, 0, 0], gx[0, :, 0]
    a00 = g[gy][:, gx]; a01 = g[gy][:, gx + 1]; a10 = g[gy + 1][:, gx]; a11 = g[gy + 1][:, gx + 1]
    top = a00 * (64 - fx) + a01 * fx
    bot = a10 * (64 - fx) + a11 * fx
    return (top * (64 - fy) + bot * fy) // 4096


def rgb(h, w, seed, channels=3):
    rng = np.random.default_rng(seed)
    img = _smooth(rng, h, w, channels)
    img = img + rng.integers(-6, 7, (h, w, 1)) + rng.integers(-3, 4, (h, w, channels))
    return np.ascontiguousarray(np.clip(img, 0, 255).astype(np.uint8))


def gray_as_rgb(h, w, seed):
    """Single-plane content replicated to R==G==B (the reference's 'grayscale' tile path)."""
    return np.ascontiguousarray(np.repeat(rgb(h, w, seed, channels=1), 3, axis=2))


def rgba(h, w, seed):
    """RGB + alpha from a 128-px block map {0, 255, ramp}; colour zeroed where alpha == 0."""
    rng = np.random.default_rng(seed)
    base = rgb(h, w, seed + 7919)
    kind = rng.integers(0, 3, (h // 128 + 1, w // 128 + 1))
    kind = np.repeat(np.repeat(kind, 128, axis=0), 128, axis=1)[:h, :w]
    ys, xs = np.mgrid[0:h, 0:w]
    ramp = ((2 * xs + ys) % 256).astype(np.uint8)
    alpha = np.where(kind == 0, 0, np.where(kind == 1, 255, ramp)).astype(np.uint8)
    out = np.concatenate([base, alpha[..., None]], axis=2)
    out[alpha == 0] = 0
    return np.ascontiguousarray(out)


def sintel_like(seed, h=1080, w=1920):
    """1080p frame with black letterbox rows [0,131) and [949,1080) (config 3)."""
    f = rgb(h, w, seed)
    f[:131] = 0
    f[949:] = 0
    return f


def noise(h, w, seed, channels=3):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (h, w, channels), dtype=np.uint8)
