"""ctypes mirror of include/xpng_b200.h (batch C ABI) and include/xpng.h / seven.h (file API).

Names and argument meaning follow the reference: `xpng_store(mode, pm, path)`, `xpng_load(path)`,
`load_7`, `store_7` (xpng.h:12-20, 7/seven.h:3-4); errors are the reference's "returns 1", surfaced
here as exceptions carrying xpngb_last_error().
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class LibraryMissing(RuntimeError):
    pass


def lib_path():
    # XPNGB_LIB: developer override for A/B builds of the same library (tools/abtest.sh)
    return os.environ.get("XPNGB_LIB") or os.path.join(HERE, "libxpng_b200.so")


class _Image(C.Structure):
    _fields_ = [("w", C.c_uint64), ("h", C.c_uint64), ("offset", C.c_uint64), ("A", C.c_uint32), ("mode", C.c_uint32)]


class _Xpng(C.Structure):  # xpng_t, xpng.h:10
    _fields_ = [("p", C.c_void_p), ("w", C.c_uint64), ("h", C.c_uint64), ("s", C.c_uint64), ("A", C.c_bool)]


_lib = None


def lib():
    """Load libxpng_b200.so; there is no fallback when it is absent."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise LibraryMissing(f"{p} not built: run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C xpng_b200`")
        L = C.CDLL(p)
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        L.xpngb_create.restype = C.c_int; L.xpngb_create.argtypes = [C.POINTER(vp), C.c_int]
        L.xpngb_destroy.restype = None; L.xpngb_destroy.argtypes = [vp]
        L.xpngb_last_error.restype = C.c_char_p; L.xpngb_last_error.argtypes = [vp]
        L.xpngb_encode_bound.restype = u64; L.xpngb_encode_bound.argtypes = [C.POINTER(_Image), u32]
        L.xpngb_encode.restype = C.c_int
        L.xpngb_encode.argtypes = [vp, C.c_int, C.POINTER(_Image), u32, vp, u64, C.c_int, vp, u64, C.c_int,
                                   C.POINTER(u64), C.POINTER(u64)]
        L.xpngb_transform.restype = C.c_int
        L.xpngb_transform.argtypes = [vp, C.c_int, C.POINTER(_Image), u32, vp, u64, C.c_int, vp, C.c_int]
        L.xpngb_encode_oriented.restype = C.c_int
        L.xpngb_encode_oriented.argtypes = [vp, C.c_int, C.c_int, C.POINTER(_Image), u32, vp, u64, C.c_int, vp, u64, C.c_int,
                                            C.POINTER(u64), C.POINTER(u64)]
        L.xpngb_peek.restype = C.c_int; L.xpngb_peek.argtypes = [vp, u64, C.POINTER(_Image)]
        L.xpngb_decode.restype = C.c_int
        L.xpngb_decode.argtypes = [vp, C.POINTER(_Image), u32, vp, u64, C.c_int, C.POINTER(u64), C.POINTER(u64), vp, u64, C.c_int]
        L.xpngb_last_kernel_ms.restype = C.c_float; L.xpngb_last_kernel_ms.argtypes = [vp]
        L.xpngb_last_launches.restype = u32; L.xpngb_last_launches.argtypes = [vp]
        L.xpngb_stream.restype = vp; L.xpngb_stream.argtypes = [vp]
        L.xpngb_profile.restype = None; L.xpngb_profile.argtypes = [vp, C.c_int]
        L.xpngb_profile_report.restype = u32; L.xpngb_profile_report.argtypes = [vp, C.c_char_p, u32]
        L.xpngb_ycocg_forward.restype = C.c_int; L.xpngb_ycocg_forward.argtypes = [vp, vp, vp, u64]
        L.xpngb_ycocg_inverse.restype = C.c_int; L.xpngb_ycocg_inverse.argtypes = [vp, vp, vp, u64]
        # frame batches sharded over devices / ranks (host C, no NCCL): include/xpng_b200.h
        L.xpngb_shard_range.restype = None; L.xpngb_shard_range.argtypes = [u32, u32, u32, C.POINTER(u32), C.POINTER(u32)]
        L.xpngb_packed_offsets.restype = None
        L.xpngb_packed_offsets.argtypes = [C.POINTER(u64), u32, C.POINTER(u64), C.POINTER(u64)]
        L.xpngb_pool_create.restype = C.c_int; L.xpngb_pool_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), u32]
        L.xpngb_pool_destroy.restype = None; L.xpngb_pool_destroy.argtypes = [vp]
        L.xpngb_pool_size.restype = u32; L.xpngb_pool_size.argtypes = [vp]
        L.xpngb_pool_last_error.restype = C.c_char_p; L.xpngb_pool_last_error.argtypes = [vp]
        L.xpngb_pool_encode.restype = C.c_int
        L.xpngb_pool_encode.argtypes = [vp, C.c_int, C.POINTER(_Image), u32, vp, u64, vp, u64, C.POINTER(u64), C.POINTER(u64)]
        L.xpngb_pool_decode.restype = C.c_int
        L.xpngb_pool_decode.argtypes = [vp, C.POINTER(_Image), u32, vp, u64, C.POINTER(u64), C.POINTER(u64), vp, u64]
        L.xpngb_gather_open.restype = C.c_int; L.xpngb_gather_open.argtypes = [C.POINTER(vp), C.c_char_p, u32, u32, u32]
        L.xpngb_gather_sizes.restype = C.c_int
        L.xpngb_gather_sizes.argtypes = [vp, u32, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
        L.xpngb_gather_close.restype = None; L.xpngb_gather_close.argtypes = [vp]
        for name in ("xpng_store", "xpng_load", "xpng_from_jpg", "xpng_store_T", "xpng_load_T", "xpng_from_jpg_T", "store_7", "load_7"):
            getattr(L, name).restype = C.c_bool
        L.xpng_store.argtypes = [u64, C.POINTER(_Xpng), C.c_char_p]
        L.xpng_store_T.argtypes = [u64, u64, C.POINTER(_Xpng), C.c_char_p]
        L.xpng_load.argtypes = [C.c_char_p, C.POINTER(_Xpng)]
        L.xpng_load_T.argtypes = [u64, C.c_char_p, C.POINTER(_Xpng)]
        L.xpng_from_jpg.argtypes = [C.c_char_p, C.c_char_p]
        L.store_7.argtypes = [C.POINTER(_Xpng), C.c_char_p]
        L.load_7.argtypes = [C.c_char_p, C.POINTER(_Xpng)]
        _lib = L
    return _lib


OPS = {"r90": 0, "r270": 1, "mv": 2, "mh": 3, "mvh": 4, "tl": 5, "tr": 6}   # include/xpng_b200.h, order of tool.c:133


def _align16(v):
    return (v + 15) & ~15


class Codec:
    """One codec context on one CUDA device (one per process/rank)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        if lib().xpngb_create(C.byref(self._h), int(device)):
            raise RuntimeError("xpngb_create failed: no usable CUDA device (there is no CPU fallback)")
        # CUDA is initialised now: the queue count asked for in __init__.py has been read, children need not inherit it
        import xpng_b200 as _pkg
        if getattr(_pkg, "CONNECTIONS_SET_HERE", False):
            os.environ.pop("CUDA_DEVICE_MAX_CONNECTIONS", None)
            _pkg.CONNECTIONS_SET_HERE = False

    def close(self):
        if self._h:
            lib().xpngb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self):
        return lib().xpngb_last_error(self._h).decode()

    @property
    def last_kernel_ms(self):
        return float(lib().xpngb_last_kernel_ms(self._h))

    @property
    def last_launches(self):
        return int(lib().xpngb_last_launches(self._h))

    @property
    def stream(self):
        return lib().xpngb_stream(self._h)

    def profile(self, on):
        """Per-kernel CUDA-event timing (serialises launches); see profile_report()."""
        lib().xpngb_profile(self._h, 3 if on == 3 else int(bool(on)))   # 3: timeline to stderr, launches not serialised

    def profile_report(self):
        """{kernel name: (total ms, launches)} accumulated since profile(True)."""
        buf = C.create_string_buffer(1 << 16)
        lib().xpngb_profile_report(self._h, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, cnt = line.rsplit(" ", 2)
            out[name] = (float(ms), int(cnt))
        return out

    # ---------------------------------------------------------------- raw pointer API (bench / device-resident data)
    @staticmethod
    def layout(shapes):
        """[(h, w, c)] -> (descriptor array, total bytes): images packed at 16-byte aligned offsets."""
        arr = (_Image * len(shapes))()
        off = 0
        for i, (h, w, c) in enumerate(shapes):
            arr[i].w, arr[i].h, arr[i].offset, arr[i].A, arr[i].mode = w, h, off, int(c == 4), 0
            off = _align16(off + h * w * c)
        return arr, off

    def encode_raw(self, level, descs, n, pix_ptr, pix_size, pix_dev, out_ptr, out_cap, out_dev):
        offs, sizes = (C.c_uint64 * n)(), (C.c_uint64 * n)()
        if lib().xpngb_encode(self._h, int(level), descs, n, C.c_void_p(pix_ptr), pix_size, int(pix_dev),
                              C.c_void_p(out_ptr), out_cap, int(out_dev), offs, sizes):
            raise RuntimeError("xpngb_encode: " + self._err())
        return offs, sizes

    def decode_raw(self, descs, n, files_ptr, files_size, files_dev, offs, sizes, pix_ptr, pix_cap, pix_dev):
        if lib().xpngb_decode(self._h, descs, n, C.c_void_p(files_ptr), files_size, int(files_dev), offs, sizes,
                              C.c_void_p(pix_ptr), pix_cap, int(pix_dev)):
            raise RuntimeError("xpngb_decode: " + self._err())

    # ---------------------------------------------------------------- numpy convenience (host buffers)
    def encode(self, level, images):
        """list of (h,w,3|4) uint8 arrays -> list of .xpng byte strings."""
        images = [np.ascontiguousarray(a, dtype=np.uint8) for a in images]
        descs, total = self.layout([a.shape for a in images])
        buf = np.zeros(total + 16, dtype=np.uint8)
        for d, a in zip(descs, images):
            buf[d.offset: d.offset + a.size] = a.reshape(-1)
        cap = int(lib().xpngb_encode_bound(descs, len(images)))
        out = np.empty(cap + 16, dtype=np.uint8)
        offs, sizes = self.encode_raw(level, descs, len(images), buf.ctypes.data, total, 0, out.ctypes.data, cap, 0)
        return [out[offs[i]: offs[i] + sizes[i]].tobytes() for i in range(len(images))]

    def decode(self, files):
        """list of .xpng byte strings -> list of (h,w,3|4) uint8 arrays."""
        n = len(files)
        descs = (_Image * n)()
        foffs, fsizes = (C.c_uint64 * n)(), (C.c_uint64 * n)()
        off = poff = 0
        for i, f in enumerate(files):
            b = np.frombuffer(f, dtype=np.uint8)
            if lib().xpngb_peek(b.ctypes.data, len(f), C.byref(descs[i])):
                raise ValueError(f"file {i}: not an .xpng header")
            foffs[i], fsizes[i] = off, len(f)
            off = _align16(off + len(f))
            descs[i].offset = poff
            poff = _align16(poff + descs[i].w * descs[i].h * (3 + descs[i].A))
        fb = np.zeros(off + 16, dtype=np.uint8)
        for i, f in enumerate(files):
            fb[foffs[i]: foffs[i] + len(f)] = np.frombuffer(f, dtype=np.uint8)
        px = np.empty(poff + 16, dtype=np.uint8)
        self.decode_raw(descs, n, fb.ctypes.data, off, 0, foffs, fsizes, px.ctypes.data, poff, 0)
        out = []
        for d in descs:
            c = 3 + d.A
            out.append(px[d.offset: d.offset + d.w * d.h * c].reshape(d.h, d.w, c).copy())
        return out

    # ---------------------------------------------------------------- traversal-order operations (Mirroring_and_Rotating/tool.c)
    def _pack(self, images):
        images = [np.ascontiguousarray(a, dtype=np.uint8) for a in images]
        descs, total = self.layout([a.shape for a in images])
        buf = np.zeros(total + 16, dtype=np.uint8)
        for d, a in zip(descs, images):
            buf[d.offset: d.offset + a.size] = a.reshape(-1)
        return images, descs, total, buf

    def transform(self, op, images):
        """Apply op ('r90', 'r270', 'mv', 'mh', 'mvh', 'tl', 'tr' — tool.c:133) to a list of (h,w,3|4) arrays."""
        images, descs, total, buf = self._pack(images)
        out = np.empty(total + 16, dtype=np.uint8)
        if lib().xpngb_transform(self._h, OPS[op], descs, len(images), buf.ctypes.data, total, 0, out.ctypes.data, 0):
            raise RuntimeError("xpngb_transform: " + self._err())
        return [out[d.offset: d.offset + d.w * d.h * (3 + d.A)].reshape(d.h, d.w, 3 + d.A).copy() for d in descs]

    def encode_oriented(self, level, op, images):
        """.xpng files of the images as `op` would leave them (tool --op followed by xpng -level), pixels untouched."""
        images, descs, total, buf = self._pack(images)
        n = len(images)
        cap = int(lib().xpngb_encode_bound(descs, n))
        out = np.empty(cap + 16, dtype=np.uint8)
        offs, sizes = (C.c_uint64 * n)(), (C.c_uint64 * n)()
        if lib().xpngb_encode_oriented(self._h, int(level), OPS[op], descs, n, buf.ctypes.data, total, 0, out.ctypes.data, cap, 0,
                                       offs, sizes):
            raise RuntimeError("xpngb_encode_oriented: " + self._err())
        return [out[offs[i]: offs[i] + sizes[i]].tobytes() for i in range(n)]

    def ycocg_forward(self, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
        out = np.empty((len(rgb), 3), dtype=np.int16)
        if lib().xpngb_ycocg_forward(self._h, rgb.ctypes.data, out.ctypes.data, len(rgb)):
            raise RuntimeError(self._err())
        return out

    def ycocg_inverse(self, ycc):
        ycc = np.ascontiguousarray(ycc, dtype=np.int16).reshape(-1, 3)
        out = np.empty((len(ycc), 3), dtype=np.uint8)
        if lib().xpngb_ycocg_inverse(self._h, ycc.ctypes.data, out.ctypes.data, len(ycc)):
            raise RuntimeError(self._err())
        return out


# ---------------------------------------------------------------- reference-shaped file API (xpng.h / seven.h)

def _as_xpng(px):
    px = np.ascontiguousarray(px, dtype=np.uint8)
    h, w, c = px.shape
    return _Xpng(px.ctypes.data, w, h, px.size, c == 4), px


def xpng_store(mode, px, path):
    """xpng_store(mode, pm, path) of xpng.h:12 — returns False on success (the reference's 0)."""
    pm, keep = _as_xpng(px)
    return bool(lib().xpng_store(int(mode), C.byref(pm), os.fsencode(path)))


def _take(pm):
    c = 3 + int(pm.A)
    arr = np.ctypeslib.as_array(C.cast(pm.p, C.POINTER(C.c_uint8)), shape=(pm.s,)).reshape(pm.h, pm.w, c).copy()
    C.CDLL(None).free(C.c_void_p(pm.p))
    return arr


def xpng_load(path):
    """xpng_load(path, &pm) of xpng.h:13 — returns the pixels or raises."""
    pm = _Xpng()
    if lib().xpng_load(os.fsencode(path), C.byref(pm)):
        raise RuntimeError("xpng_load failed")
    return _take(pm)


def store_7(px, path):
    pm, keep = _as_xpng(px)
    return bool(lib().store_7(C.byref(pm), os.fsencode(path)))


def load_7(path):
    pm = _Xpng()
    if lib().load_7(os.fsencode(path), C.byref(pm)):
        raise RuntimeError("load_7 failed")
    return _take(pm)
