"""xpng_b200 — B200-native (sm_100a) implementation of the xPNG encode/decode hot path.

The product is the shared library `libxpng_b200.so` (hand-written CUDA kernels behind a C ABI,
include/xpng_b200.h, plus the reference-compatible C file API of include/xpng.h / include/seven.h).
This package is the thin ctypes mirror used by the tests and bench.py; it never falls back to a CPU
implementation: importing `codec` without the built library, or creating a Codec without a CUDA
device, raises.
"""
from .codec import Codec, LibraryMissing, lib, lib_path, xpng_store, xpng_load, load_7, store_7  # noqa: F401
