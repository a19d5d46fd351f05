#!/usr/bin/env python
"""The reference's round-trip script (test.rb:1-45) re-expressed for this repo's binaries (ruby is not in the image):
for every PNG of a directory: seven --to_7, then for each level 1, 2, 7: xpng -<level>, xpng -d, cmp; prints OK / Failed and
the compressed size like test.rb:5-11.  JPEGs go through `xpng -3`, which is the reference's unimplemented stub (Failed).
Usage: python tools/test_rb.py [image_dir]      (default: tests/golden/corpus; needs a GPU)"""
import filecmp, glob, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEVEN, XPNG = os.path.join(ROOT, "xpng_b200", "bin", "seven"), os.path.join(ROOT, "xpng_b200", "bin", "xpng")

def run(*a):
    return subprocess.run(a, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 0

def main():
    d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "corpus")
    tmp = tempfile.mkdtemp(prefix="xpng_testrb_")
    failed = 0
    for path in sorted(glob.glob(os.path.join(d, "*"))):
        name = os.path.basename(path)
        if name.lower().endswith((".jpg", ".jpeg")):
            ok = run(XPNG, "-3", path, os.path.join(tmp, "j.xpng"))
            print(f"\n{name}\n\n    -3: {'OK' if ok else 'Failed'}"); continue
        if not name.lower().endswith(".png"): continue
        s7 = os.path.join(tmp, "a.7")
        print(f"\n{name}\n")
        if not run(SEVEN, "--to_7", path, s7):
            print("    PNG -> .7: Failed"); failed += 1; continue
        for o in (1, 2, 7):
            x, b = os.path.join(tmp, "a.xpng"), os.path.join(tmp, "b.7")
            ok = run(XPNG, f"-{o}", s7, x) and run(XPNG, "-d", x, b) and filecmp.cmp(s7, b, shallow=False)
            size = os.path.getsize(x) if os.path.exists(x) else 0
            print(f"    -{o}: {'OK' if ok else 'Failed':6s} {f'{size:,}'.replace(',', '_'):>12s} B")
            failed += not ok
    print(f"\n{'all round trips OK' if not failed else str(failed) + ' FAILED'}\n")
    return 1 if failed else 0

if __name__ == "__main__":
    sys.exit(main())
