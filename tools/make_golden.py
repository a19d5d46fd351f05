#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ with the UNMODIFIED reference (oracle/_ref/xpng,
built from /root/reference by oracle/Makefile).  Run in the build container (needs the reference
tree for the corpus and the _ref binaries); the outputs are committed:

  tests/golden/crops/<name>.png   small crops of the reference's images/*.png corpus (test inputs)
  tests/golden/manifest.json      per case: shape, sha256 + size of the reference .xpng per level,
                                  sha256 of the normalised pixels the reference decodes back

Full-size corpus PNGs are copied to tests/golden/corpus/ (committed: the config-0 inputs).
"""
import hashlib, json, os, shutil, sys
import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po          # noqa: E402
from xpng_b200 import synth                # noqa: E402

REF_IMAGES = "/root/reference/images"
GOLD = os.path.join(ROOT, "tests", "golden")


def load_png(path):
    im = Image.open(path)
    im = im.convert("RGBA" if ("A" in im.getbands() or "transparency" in im.info) else "RGB")
    return np.ascontiguousarray(np.asarray(im, dtype=np.uint8))


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def entry(px):
    e = {"shape": list(px.shape), "levels": {}}
    for lv in (1, 2, 7):
        f = po.ref_encode(lv, px)
        assert f is not None
        back = po.ref_decode(f)
        e["levels"][str(lv)] = {"size": len(f), "sha256": sha(f)}
        e["decoded_sha256"] = sha(back.tobytes())
        e["decoded_shape"] = list(back.shape)
    return e


def synthetic_cases(full):
    c = {
        "rgb_600x500_s1": ("rgb", [500, 600, 1]),
        "rgb_1334x750_s11": ("rgb", [750, 1334, 11]),
        "rgba_700x520_s2": ("rgba", [520, 700, 2]),
        "gray_512x512_s3": ("gray_as_rgb", [512, 512, 3]),
        "noise_500x500_s4": ("noise", [500, 500, 4]),
        "rgb_37x1500_s5": ("rgb", [1500, 37, 5]),
        "rgb_1500x37_s6": ("rgb", [37, 1500, 6]),
        "rgb_3x3_s9": ("rgb", [3, 3, 9]),
        "rgb_2x1_s8": ("rgb", [1, 2, 8]),
        "rgb_1x1_s7": ("rgb", [1, 1, 7]),
        "rgb_445x445_s12": ("rgb", [445, 445, 12]),
        "rgb_667x889_s13": ("rgb", [889, 667, 13]),
        "rgba_5x4_s14": ("rgba", [4, 5, 14]),
        "sintel_1080p_s1000": ("sintel_like", [1000]),
        "sintel_1080p_s1001": ("sintel_like", [1001]),
    }
    if full:
        c.update({
            "rgb_4k_s1": ("rgb", [2160, 3840, 1]),
            "gray_4096_s3000": ("gray_as_rgb", [4096, 4096, 3000]),
            "rgba_8192_s2": ("rgba", [8192, 8192, 2]),
        })
    return c


def main():
    assert po.ref_available(), "build oracle/_ref first (make -C oracle)"
    full = "--full" in sys.argv
    man = {"synthetic": {}, "crops": {}, "corpus": {}, "special": {}}
    for name, (fn, args) in synthetic_cases(full).items():
        px = getattr(synth, fn)(*args)
        e = entry(px); e["gen"] = [fn, args]
        man["synthetic"][name] = e
        print(name, {k: v["size"] for k, v in e["levels"].items()}, flush=True)
    # special hand-made cases
    flat = np.full((100, 100, 3), 77, np.uint8)
    man["special"]["flat_rgb_100"] = dict(entry(flat), gen=["full", [[100, 100, 3], 77]])
    flat4 = np.full((100, 100, 4), 77, np.uint8)
    man["special"]["flat_rgba_100"] = dict(entry(flat4), gen=["full", [[100, 100, 4], 77]])
    opaque = np.concatenate([synth.rgb(90, 120, 21), np.full((90, 120, 1), 255, np.uint8)], axis=2)
    man["special"]["opaque_rgba_120x90"] = dict(entry(opaque), gen=["opaque", [90, 120, 21]])
    dirty = synth.rgba(200, 300, 22); dirty[dirty[..., 3] == 0] = [9, 8, 7, 0]
    man["special"]["dirty_rgba_300x200"] = dict(entry(dirty), gen=["dirty", [200, 300, 22]])
    halfflat = synth.rgb(900, 900, 23); halfflat[:, :450] = [10, 200, 30]
    man["special"]["halfflat_900"] = dict(entry(halfflat), gen=["halfflat", [900, 900, 23]])
    os.makedirs(os.path.join(GOLD, "crops"), exist_ok=True)
    corpus_dir = os.path.join(GOLD, "corpus"); os.makedirs(corpus_dir, exist_ok=True)
    for fn in sorted(os.listdir(REF_IMAGES)):
        if not fn.endswith(".png"):
            continue
        px = load_png(os.path.join(REF_IMAGES, fn))
        shutil.copy(os.path.join(REF_IMAGES, fn), os.path.join(corpus_dir, fn))
        man["corpus"][fn] = entry(px)
        h, w, _ = px.shape
        ch, cw = min(h, 160), min(w, 224)
        y0, x0 = (h - ch) // 2, (w - cw) // 2
        crop = np.ascontiguousarray(px[y0:y0 + ch, x0:x0 + cw])
        Image.fromarray(crop).save(os.path.join(GOLD, "crops", fn), optimize=True)
        assert np.array_equal(load_png(os.path.join(GOLD, "crops", fn)), crop) or crop.shape[2] == 4
        crop = load_png(os.path.join(GOLD, "crops", fn))
        man["crops"][fn] = entry(crop)
        print(fn, px.shape, {k: v["size"] for k, v in man["corpus"][fn]["levels"].items()}, flush=True)
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
