#!/usr/bin/env python
"""BASELINE config 3 on N GPUs: F sintel-like 1080p frames (seeds 1000+i) cut into contiguous shards,
every rank codes its shard device-resident with ONE batch call per direction, the only exchange is the all-gather of
the compressed sizes (xpng_b200.shard.global_table).  Times are CUDA-event kernel times, MAX over ranks.
  python tools/shard_bench.py --frames 128                                   (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
         tools/shard_bench.py --frames 128
The printed `sizes_sha` must not depend on N: the files of a frame are the same whichever rank coded it."""
import argparse, hashlib, json, os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xpng_b200
from xpng_b200 import synth, Codec
from xpng_b200.shard import shard_range, global_table

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=128)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
lo, hi = shard_range(args.frames, rank, world)
imgs = [synth.sintel_like(1000 + i) for i in range(lo, hi)]
n = len(imgs)
lib = xpng_b200.lib()
cd = Codec(local)
shapes = [a.shape for a in imgs]
descs, total = Codec.layout(shapes)
buf = np.zeros(total + 64, np.uint8)
for d, a in zip(descs, imgs): buf[d.offset:d.offset + a.size] = a.reshape(-1)
cap = int(lib.xpngb_encode_bound(descs, n))
d_px = torch.from_numpy(buf).to(dev); d_f = torch.zeros(cap + 64, dtype=torch.uint8, device=dev); d_back = torch.zeros(total + 64, dtype=torch.uint8, device=dev)
npx_all = args.frames * 1920 * 1080 / 1e6
out = {"frames": args.frames, "n_gpus": world, "frames_per_rank": n}
for lv in (1, 2):
    be = bd = 1e9
    for _ in range(args.reps):
        if world > 1: dist.barrier()
        d, _ = Codec.layout(shapes)
        offs, sz = cd.encode_raw(lv, d, n, d_px.data_ptr(), total, 1, d_f.data_ptr(), cap, 1)
        e = cd.last_kernel_ms
        table_off, table_sz = global_table([int(s) for s in sz], args.frames)     # the one exchange: sizes -> global offsets
        d2, _ = Codec.layout(shapes)
        for x in d2: x.w = x.h = 0
        d_back.zero_()
        cd.decode_raw(d2, n, d_f.data_ptr(), cap, 1, offs, sz, d_back.data_ptr(), total, 1)
        k = cd.last_kernel_ms
        t = torch.tensor([e, k], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        be, bd = min(be, float(t[0])), min(bd, float(t[1]))
    ok = torch.tensor([int(torch.equal(d_back[:total], d_px[:total]))], device=dev)
    if world > 1: dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out[f"L{lv}"] = {"enc_ms": round(be, 2), "dec_ms": round(bd, 2), "enc_MPix_s": round(npx_all / be * 1e3), "dec_MPix_s": round(npx_all / bd * 1e3),
                     "xpng_bytes": int(sum(table_sz)), "arena_bytes": int(table_off[-1] + table_sz[-1]),
                     "sizes_sha": hashlib.sha256(np.array(table_sz, dtype=np.int64).tobytes()).hexdigest()[:16], "roundtrip": bool(ok.item())}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
