#!/usr/bin/env python
"""Host<->device copy bandwidth of the box with pinned memory: each direction alone, then both at once on two streams.
The e2e leg of bench.py moves 35 GB per step; this is its floor.  python tools/pcie_probe.py [GiB]"""
import sys, time, torch
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gib * (1 << 30))
h_a = torch.empty(n, dtype=torch.uint8).pin_memory(); h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if up:
            with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
for name, up, down in (("H2D alone", 1, 0), ("D2H alone", 0, 1), ("H2D + D2H at once", 1, 1)):
    t = run(up, down)
    print(f"{name:20s} {gib * (up + down) / t * 1.0737:7.1f} GB/s total ({t * 1e3:.1f} ms for {gib * (up + down):.0f} GiB)")
