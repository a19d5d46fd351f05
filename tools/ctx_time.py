#!/usr/bin/env python
"""Process start-up cost of the CLI against CUDA_DEVICE_MAX_CONNECTIONS (the batch path wants 32 hardware queues)."""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xpng_b200 import synth
a = synth.rgb(256, 256, 1)
open("/tmp/t.7", "wb").write(np.array([(256 - 1) + (7 << 24), (256 - 1)], dtype=np.uint32).tobytes() + a.tobytes())
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "xpng_b200", "bin", "xpng")
for c in ("8", "16", "32"):
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS=c)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); subprocess.run([exe, "-1", "/tmp/t.7", "/tmp/t.xpng"], env=env, stdout=subprocess.DEVNULL, check=True); ts.append(time.perf_counter() - t0)
    print(f"CUDA_DEVICE_MAX_CONNECTIONS={c}: xpng -1 of a 256x256 image takes {min(ts):.2f} s (process start to exit, best of 3)")
