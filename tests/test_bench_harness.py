"""Host-side pieces of bench.py that can be checked without a GPU: the file-API harness (against the reference built by
oracle/Makefile, which exports the same xpng.h entry points as the drop-in) and the NUMA binding helper."""
import ctypes as C
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from xpng_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "libxpng_ref.so")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("callers", [1, 3])
def test_file_api_steps_one_and_many_callers(callers):
    """Every frame is stored and loaded at levels 1 and 2, from one caller or several; no file is left behind."""
    frames = [synth.sintel_like(1000 + k, 270, 480) for k in range(3)]
    per_step, calls = bench.file_api_steps(C.CDLL(REF), frames, 1, 1, f"t{callers}", callers=callers)
    assert per_step > 0 and set(calls) == {"enc1", "dec1", "enc2", "dec2"} and all(v > 0 for v in calls.values())
    assert not [f for f in os.listdir(bench._tmpdir()) if f.startswith(f"_xpng_t{callers}_{os.getpid()}_")]


def test_bind_near_gpu_never_raises_and_keeps_the_affinity_without_a_gpu():
    before = os.sched_getaffinity(0)
    rec, orig = bench.bind_near_gpu(0)
    try:
        assert orig == before and isinstance(rec, dict) and "numa_node" in rec
        assert os.sched_getaffinity(0) <= before
    finally:
        os.sched_setaffinity(0, before)
