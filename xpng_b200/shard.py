"""Frame batches sharded over the GPUs of one box, SURVEY.md §8(e): a thin ctypes mirror of the C entry points of
include/xpng_b200.h (xpng_b200/host/xpng_pool.c).  No torch, no NCCL: the only exchange between shards is the table
of compressed sizes, and it travels either inside one process (Pool: one host thread and one codec context per
device) or through a POSIX shared-memory segment (Gather: one process per device, torchrun-style launches).

Frames are independent, so a batch is cut into contiguous shards (frame i -> shard floor(i * world / n)) and every
shard is coded with no data-path collective; every participant derives the same global offset table from the gathered
sizes by an exclusive scan (files at 16-byte aligned offsets, as `xpngb_encode` lays them out inside one arena)."""
import ctypes as C

import numpy as np

from .codec import _Image, _align16, lib


def shard_range(n_frames, rank, world):
    """Half-open range of the frames shard `rank` codes: frame i belongs to shard floor(i * world / n_frames)."""
    first, count = C.c_uint32(), C.c_uint32()
    lib().xpngb_shard_range(int(n_frames), int(world), int(rank), C.byref(first), C.byref(count))
    return first.value, first.value + count.value


def owner(i, n_frames, world):
    return i * world // n_frames


def packed_offsets(sizes):
    """Exclusive scan of the 16-byte padded sizes: (offsets, total bytes)."""
    n = len(sizes)
    a = (C.c_uint64 * n)(*[int(s) for s in sizes]); o = (C.c_uint64 * n)(); tot = C.c_uint64()
    lib().xpngb_packed_offsets(a, n, o, C.byref(tot))
    return list(o), tot.value


class Gather:
    """The size/offset gather between the ranks of one box (one process per GPU)."""

    def __init__(self, name, rank, world, max_items):
        self._h = C.c_void_p()
        self.rank, self.world = int(rank), int(world)
        if lib().xpngb_gather_open(C.byref(self._h), str(name).encode(), self.rank, self.world, int(max_items)):
            raise RuntimeError("xpngb_gather_open failed")

    def sizes(self, n_frames, local_sizes):
        """Collective: returns (offsets, sizes) of the whole batch on every rank."""
        lo, hi = shard_range(n_frames, self.rank, self.world)
        assert len(local_sizes) == hi - lo, "one size per frame of the local shard"
        loc = (C.c_uint64 * max(1, hi - lo))(*[int(s) for s in local_sizes])
        all_s, all_o = (C.c_uint64 * n_frames)(), (C.c_uint64 * n_frames)()
        if lib().xpngb_gather_sizes(self._h, int(n_frames), loc, all_s, all_o):
            raise RuntimeError("xpngb_gather_sizes failed (timeout or size mismatch)")
        return list(all_o), list(all_s)

    def close(self):
        if self._h:
            lib().xpngb_gather_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pool:
    """One process driving several devices: xpngb_pool_encode / xpngb_pool_decode with host buffers."""

    def __init__(self, devices):
        devices = list(devices)
        self._h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        if lib().xpngb_pool_create(C.byref(self._h), arr, len(devices)):
            raise RuntimeError("xpngb_pool_create failed: no usable CUDA devices")

    def close(self):
        if self._h:
            lib().xpngb_pool_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self):
        return lib().xpngb_pool_last_error(self._h).decode()

    def encode(self, level, images):
        """list of (h,w,3|4) uint8 arrays -> (files, offsets, sizes): the files and the global table of the sharded call."""
        images = [np.ascontiguousarray(a, dtype=np.uint8) for a in images]
        n = len(images)
        descs = (_Image * n)()
        off = 0
        for i, a in enumerate(images):
            h, w, c = a.shape
            descs[i].w, descs[i].h, descs[i].offset, descs[i].A, descs[i].mode = w, h, off, int(c == 4), 0
            off = _align16(off + a.size)
        buf = np.zeros(off + 16, dtype=np.uint8)
        for d, a in zip(descs, images):
            buf[d.offset: d.offset + a.size] = a.reshape(-1)
        cap = int(lib().xpngb_encode_bound(descs, n))
        out = np.empty(cap + 16, dtype=np.uint8)
        offs, sizes = (C.c_uint64 * n)(), (C.c_uint64 * n)()
        if lib().xpngb_pool_encode(self._h, int(level), descs, n, buf.ctypes.data, off, out.ctypes.data, cap, offs, sizes):
            raise RuntimeError("xpngb_pool_encode: " + self._err())
        return [out[offs[i]: offs[i] + sizes[i]].tobytes() for i in range(n)], list(offs), list(sizes)

    def decode(self, files):
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
        if lib().xpngb_pool_decode(self._h, descs, n, fb.ctypes.data, off, foffs, fsizes, px.ctypes.data, poff):
            raise RuntimeError("xpngb_pool_decode: " + self._err())
        return [px[d.offset: d.offset + d.w * d.h * (3 + d.A)].reshape(d.h, d.w, 3 + d.A).copy() for d in descs]
