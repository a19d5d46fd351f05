"""PNG <-> .7 converter (xpng_b200/bin/seven, host C on zlib; reference 7/seven.c on libpng).  PIL is the checker:
the reference's tool is `png_image_finish_read` to RGB / RGBA followed by normalize_RGBA (7/seven.c:4-37, :39-65)."""
import glob
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

from conftest import ROOT
from oracle import pyoracle as po

SEVEN = os.path.join(ROOT, "xpng_b200", "bin", "seven")
XPNG = os.path.join(ROOT, "xpng_b200", "bin", "xpng")


def _expected(png_path):
    ref = Image.open(png_path)
    ref = ref.convert("RGBA" if (ref.mode in ("RGBA", "LA") or "transparency" in ref.info) else "RGB")
    return po.normalize(np.ascontiguousarray(np.array(ref)))


def _variants():
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (41, 67, 3), dtype=np.uint8)
    rgba = rng.integers(0, 256, (41, 67, 4), dtype=np.uint8)
    opaque = rgba.copy(); opaque[..., 3] = 255
    dirty = rgba.copy(); dirty[::3, ::2, 3] = 0
    out = {"rgb": Image.fromarray(rgb), "rgba": Image.fromarray(rgba), "rgba_opaque": Image.fromarray(opaque),
           "rgba_dirty": Image.fromarray(dirty), "grey8": Image.fromarray(rgb[..., 0]), "grey1": Image.fromarray(rgb[..., 0] > 99),
           "grey_alpha": Image.fromarray(rgba[..., :2].copy(), "LA"), "palette": Image.fromarray(rgb).convert("P"),
           "palette16": Image.fromarray(rgb).quantize(16), "palette2": Image.fromarray(rgb).quantize(2), "one_pixel": Image.fromarray(rgb[:1, :1])}
    return out


@pytest.mark.parametrize("name", sorted(_variants()))
def test_to_7_matches_pil(tmp_path, name):
    im = _variants()[name]
    png, seven = str(tmp_path / "a.png"), str(tmp_path / "a.7")
    im.save(png, bits=4) if name == "palette16" else (im.save(png, bits=1) if name == "palette2" else im.save(png))
    assert subprocess.run([SEVEN, "--to_7", png, seven]).returncode == 0
    want = _expected(png)
    got = po.read_7(seven)
    assert got.shape == want.shape and np.array_equal(got, want)
    back = str(tmp_path / "b.png")
    assert subprocess.run([SEVEN, "--to_png", seven, back]).returncode == 0
    assert np.array_equal(np.array(Image.open(back)), got)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "crops", "*.png"))) +
                         sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "corpus", "*.png"))), ids=os.path.basename)
def test_reference_images(tmp_path, path):
    seven = str(tmp_path / "a.7")
    assert subprocess.run([SEVEN, "--to_7", path, seven]).returncode == 0
    want = _expected(path)
    got = po.read_7(seven)
    assert got.shape == want.shape and np.array_equal(got, want)


def _adam7_png(a):
    """Hand-made Adam7 file (PIL cannot write interlaced PNGs): the seven passes, filter 0, one IDAT."""
    import struct
    import zlib

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    h, w, c = a.shape
    xs, ys, dx, dy = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1], [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
    raw = b""
    for p in range(7):
        sub = a[ys[p]::dy[p], xs[p]::dx[p]]
        for row in sub if sub.size else []:
            raw += b"\0" + row.tobytes()
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, {3: 2, 4: 6}[c], 0, 0, 1)) +
            chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


@pytest.mark.parametrize("shape", [(37, 53, 3), (1, 1, 3), (5, 3, 4), (64, 64, 4), (9, 2, 3)])
def test_adam7_interlaced(tmp_path, shape):
    a = np.random.default_rng(3).integers(0, 256, shape, dtype=np.uint8)
    png, seven = str(tmp_path / "i.png"), str(tmp_path / "i.7")
    open(png, "wb").write(_adam7_png(a))
    assert np.array_equal(np.array(Image.open(png)), a)            # PIL agrees the file is a valid interlaced PNG
    assert subprocess.run([SEVEN, "--to_7", png, seven]).returncode == 0
    want = po.normalize(a)
    got = po.read_7(seven)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_rejections_and_usage(tmp_path):
    a16 = (np.arange(40 * 30, dtype=np.uint16).reshape(30, 40) * 50)
    p16 = str(tmp_path / "g16.png"); Image.fromarray(a16).save(p16)
    assert subprocess.run([SEVEN, "--to_7", p16, str(tmp_path / "x.7")]).returncode == 1          # 7/seven.c:48
    good = str(tmp_path / "ok.png"); Image.fromarray(np.zeros((9, 9, 3), np.uint8) + 7).save(good)
    raw = bytearray(open(good, "rb").read()); raw[-20] ^= 0x55
    bad = str(tmp_path / "bad.png"); open(bad, "wb").write(raw)
    assert subprocess.run([SEVEN, "--to_7", bad, str(tmp_path / "y.7")]).returncode == 1          # CRC / stream error
    assert subprocess.run([SEVEN, "--to_7", str(tmp_path / "missing.png"), str(tmp_path / "z.7")]).returncode == 1
    r = subprocess.run([SEVEN], capture_output=True, text=True)
    assert r.returncode == 1 and "--to_7" in r.stdout and "--to_png" in r.stdout                   # 7/seven.c:73-78


@pytest.mark.gpu
def test_exchange_path_end_to_end(tmp_path):
    """test.rb of the reference, re-expressed: PNG -> .7 -> .xpng (levels 1, 2, 7) -> .7 (cmp) -> PNG."""
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "crops", "*.png")))[:5]:
        src = str(tmp_path / "src.7")
        assert subprocess.run([SEVEN, "--to_7", path, src]).returncode == 0
        px = po.read_7(src)
        for lv in (1, 2, 7):
            xp, back = str(tmp_path / "r.xpng"), str(tmp_path / "r.7")
            assert subprocess.run([XPNG, f"-{lv}", src, xp], capture_output=True).returncode == 0
            assert open(xp, "rb").read() == po.encode(lv, px)
            assert subprocess.run([XPNG, "-d", xp, back], capture_output=True).returncode == 0
            assert open(back, "rb").read() == open(src, "rb").read()                               # test.rb:32 `cmp`
        png = str(tmp_path / "out.png")
        assert subprocess.run([SEVEN, "--to_png", back, png]).returncode == 0
        assert np.array_equal(np.array(Image.open(png)), px)


# ---------------------------------------------------------------- P6 front end (ancestor/gray.c:667-682)

def test_ppm_round_trip_and_header(tmp_path):
    rgb = np.random.default_rng(9).integers(0, 256, (37, 53, 3), dtype=np.uint8)
    ppm, seven, back = str(tmp_path / "a.ppm"), str(tmp_path / "a.7"), str(tmp_path / "b.ppm")
    Image.fromarray(rgb).save(ppm)                                  # PIL writes "P6\n53 37\n255\n"
    assert subprocess.run([SEVEN, "--to_7", ppm, seven]).returncode == 0
    assert np.array_equal(po.read_7(seven), rgb)
    assert subprocess.run([SEVEN, "--to_ppm", seven, back]).returncode == 0
    raw = open(back, "rb").read()
    assert raw == b"P6\n53 37\n255\n" + rgb.tobytes()              # the header line of ancestor/gray.c:680
    assert np.array_equal(np.array(Image.open(back)), rgb)


def test_ppm_header_forms_and_rejections(tmp_path):
    rgb = np.random.default_rng(10).integers(0, 256, (3, 5, 3), dtype=np.uint8)
    seven = str(tmp_path / "a.7")
    good = [b"P6 5 3 255\n", b"P6\n# made by hand\n5\t3\n# another\n255\n", b"P6\r\n5 3\r\n255 "]
    for i, hdr in enumerate(good):
        p = tmp_path / ("g%d.ppm" % i)
        p.write_bytes(hdr + rgb.tobytes())
        assert subprocess.run([SEVEN, "--to_7", str(p), seven]).returncode == 0, hdr
        assert np.array_equal(po.read_7(seven), rgb), hdr
    bad = [b"P6\n5 3\n65535\n" + bytes(90), b"P6\n5 3\n255\n" + rgb.tobytes()[:-1], b"P6\n0 3\n255\n", b"P6\n5 3\n", b"P6\n5 x\n255\n" + rgb.tobytes(),
           b"P5\n5 3\n255\n" + rgb.tobytes(), b"P6\n5 3\n255" + rgb.tobytes()]
    for i, data in enumerate(bad):
        p = tmp_path / ("b%d.ppm" % i)
        p.write_bytes(data)
        assert subprocess.run([SEVEN, "--to_7", str(p), seven]).returncode == 1, data[:16]
    # P6 has no alpha channel: an RGBA .7 is refused
    po.write_7(seven, np.random.default_rng(1).integers(0, 256, (4, 4, 4), dtype=np.uint8))
    assert subprocess.run([SEVEN, "--to_ppm", seven, str(tmp_path / "x.ppm")]).returncode == 1
