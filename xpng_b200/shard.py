"""Frame-batch sharding over ranks (one process per GPU), SURVEY.md §8(e).

Frames are independent, so a batch is cut into contiguous shards (frame i -> rank floor(i * world / n)) and every
rank codes its shard with no data-path collective.  The only exchange is the final gather of per-frame
compressed sizes, from which every rank derives the same global offset table (exclusive scan, files at
16-byte aligned offsets like `xpngb_encode` lays them out inside one arena)."""
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world):
    """Half-open range of the frames rank `rank` codes: frame i belongs to rank floor(i * world / n_frames)."""
    lo = -(-rank * n_frames // world)          # ceil(rank * n / world)
    hi = -(-(rank + 1) * n_frames // world)
    return lo, hi


def owner(i, n_frames, world):
    return i * world // n_frames


def global_table(local_sizes, n_frames, group=None):
    """All-gather the per-frame sizes of every shard; returns (offsets, sizes) lists for the whole batch.
    Works on any backend (sizes travel as int64 tensors on the backend's device)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(n_frames, rank, world)
    assert len(local_sizes) == hi - lo, "one size per frame of the local shard"
    if world == 1:
        sizes = [int(s) for s in local_sizes]
    else:
        width = max(shard_range(n_frames, r, world)[1] - shard_range(n_frames, r, world)[0] for r in range(world))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        mine = torch.zeros(width, dtype=torch.int64, device=dev)
        if hi > lo:
            mine[: hi - lo] = torch.tensor([int(s) for s in local_sizes], dtype=torch.int64, device=dev)
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        sizes = []
        for r in range(world):
            a, b = shard_range(n_frames, r, world)
            sizes += [int(v) for v in parts[r][: b - a].tolist()]
    offsets, off = [], 0
    for s in sizes:
        offsets.append(off)
        off = (off + s + 15) & ~15
    return offsets, sizes
