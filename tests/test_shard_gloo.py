"""N > 1 host logic on CPU: two ranks (launched and fenced with gloo) shard a frame batch, code their shards (with the
oracle standing in for the GPU codec, which does not exist in this container) and agree on the global size/offset
table through the C shared-memory gather of include/xpng_b200.h (xpngb_gather_*)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as po
from xpng_b200 import shard, synth


def _frames(n):
    return [synth.rgb(40 + 3 * i, 64 + 5 * i, 500 + i) for i in range(n)]


def _worker(rank, world, port, n, level, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = _frames(n)
        lo, hi = shard.shard_range(n, rank, world)
        files = [po.encode(level, f) for f in frames[lo:hi]]
        # the size/offset gather is the C one (shared memory, no torch.distributed): gloo is only the launcher's barrier here
        g = shard.Gather(f"test{port}", rank, world, max(n, 1))
        offsets, sizes = g.sizes(n, [len(f) for f in files])
        offsets2, sizes2 = g.sizes(n, [len(f) + 1 for f in files])     # a second round through the other half of the segment
        assert sizes2 == [s + 1 for s in sizes]
        out.put((rank, lo, hi, offsets, sizes, [bytes(f) for f in files]))
        dist.barrier()
        g.close()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n,world", [(7, 2), (2, 2), (1, 2)])
def test_two_ranks_agree_on_the_table(n, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 1, q)) for port in [_free_port()] for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [po.encode(1, f) for f in _frames(n)]
    # the shards partition the batch in order
    assert [g[1] for g in got] + [n] == [shard.shard_range(n, r, world)[0] for r in range(world)] + [n]
    assert sum((g[5] for g in got), []) == want
    # every rank derived the same table, equal to the single-process layout
    offs, off = [], 0
    for f in want:
        offs.append(off)
        off = (off + len(f) + 15) & ~15
    for g in got:
        assert g[4] == [len(f) for f in want] and g[3] == offs


def test_shard_ranges_cover_and_match_owner():
    for n in (1, 2, 5, 8, 1000):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard.shard_range(n, r, world)
                seen += list(range(lo, hi))
                assert all(shard.owner(i, n, world) == r for i in range(lo, hi))
            assert seen == list(range(n))
