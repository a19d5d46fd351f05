#!/usr/bin/env python
"""bench.py — xPNG encode/decode MPix/s at -1/-2/-7 on B200 (BASELINE.json metric), one JSON line.

A "step" is one pass of the hot path over one batch: encode the workload's frame(s) at levels
1, 2 and 7 and decode each result (6 codec calls).  Workload at every N: configs[1] of BASELINE.json,
one 3840x2160 RGB synthetic frame per GPU (seed 1 + rank, SURVEY.md §8(d) generator); with N ranks
each rank codes its own frame (frames are independent: weak scaling, no data-path collective).

  value : whole-job MPix/s (pixels through the 6 calls, all ranks) with inputs and outputs resident
          in HBM, timed with CUDA events on the codec's stream, max over ranks.
  e2e   : the same step through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region).
  roofline / cpu_baseline: see DESIGN.md "Measurement".

--impl reference times the unmodified reference (oracle/_ref, built from /root/reference by
oracle/Makefile) on the host cores with all the threads it spawns by itself (T = 0 -> nproc).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
W, H = 3840, 2160
LEVELS = (1, 2, 7)
# the levels of a step are independent jobs: level 2 (the longest chains) gets a host thread of its own, levels 1 and 7
# share the second one (tools/e2e_probe.py: a third pipeline in flight only adds PCIe contention to level 2's copies)
PIPELINES = ((2,), (1, 7))
METRIC = "xpng encode+decode MPix/s over levels -1/-2/-7"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def _run_nvml(self):
        """Fast path: NVML in-process (10 ms period); returns False when NVML is unavailable."""
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            bits = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown if hasattr(N, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                    "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception:
            return False
        while not self._stop_evt.is_set():
            try:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                self.rows.append([str(self.index), str(sm), str(mx)] + ["Active" if r & bits[k] else "Not Active"
                                                                         for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            except Exception:
                pass
            self._stop_evt.wait(0.01)
        return True

    def run(self):
        if self._run_nvml():
            return
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(r[1]) for r in self.rows if r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ reference arm

def _quiet_call(fn, *a):
    """Run fn with fd 1 pointed at /dev/null (the reference prints a MPx/s line per call)."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        return fn(*a)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


class _Xpng(C.Structure):
    _fields_ = [("p", C.c_void_p), ("w", C.c_uint64), ("h", C.c_uint64), ("s", C.c_uint64), ("A", C.c_bool)]


def reference_steps(frame, steps, warmup):
    """Time the CPU baseline: K steps (encode + decode at levels 1/2/7) of the reference's own code.
    Returns (seconds per step list, kind, cores, sample)."""
    from oracle import pyoracle as po
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libxpng_ref.so")
    times = []
    if os.path.exists(ref_so):
        L = C.CDLL(ref_so)
        L.xpng_store_T.restype = C.c_bool
        L.xpng_store_T.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(_Xpng), C.c_char_p]
        L.xpng_load_T.restype = C.c_bool
        L.xpng_load_T.argtypes = [C.c_uint64, C.c_char_p, C.POINTER(_Xpng)]
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        pm = _Xpng(frame.ctypes.data, frame.shape[1], frame.shape[0], frame.size, frame.shape[2] == 4)
        paths = {lv: os.path.join(tmp, f"_xpng_ref_{os.getpid()}_{lv}.xpng").encode() for lv in LEVELS}

        def step():
            for lv in LEVELS:
                assert not L.xpng_store_T(0, lv, C.byref(pm), paths[lv])
            for lv in LEVELS:
                out = _Xpng()
                assert not L.xpng_load_T(0, paths[lv], C.byref(out))
                libc.free(C.c_void_p(out.p))
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            _quiet_call(step)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        for p in paths.values():
            if os.path.exists(p):
                os.remove(p)
        return times, "reference", os.cpu_count(), "same 3840x2160 frame, levels 1/2/7 encode+decode via xpng_store_T/xpng_load_T (T=0: all host cores), files on tmpfs"
    # the unmodified reference did not travel: time the single-threaded oracle port instead
    def step():
        for lv in LEVELS:
            f = po.encode(lv, frame)
            po.decode(f)
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, "port", 1, "same 3840x2160 frame, levels 1/2/7 encode+decode with the single-threaded oracle port"


def run_reference(args, rank, world):
    if rank != 0:
        return
    from xpng_b200 import synth
    frame = synth.rgb(H, W, 1)
    steps = max(1, min(args.steps, 20))
    times, kind, cores, sample = reference_steps(frame, steps, min(args.warmup, 2))
    per_step = sum(times) / len(times)
    value = 6 * W * H / 1e6 / per_step
    line = {"metric": METRIC, "value": round(value, 2), "unit": "MPix/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2),
            "ms_per_step": round(per_step * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "configs[1]: one 3840x2160 RGB synthetic frame, levels -1/-2/-7, encode+decode", "host_threads": cores},
            "cpu_baseline": {"value": round(value, 2), "unit": "MPix/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(value, 2), "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm

def run_ours(args, rank, world, local_rank):
    import torch
    import xpng_b200
    from xpng_b200 import synth
    from oracle import pyoracle as po

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        # the communicator is created here (eagerly, or by the barrier): NCCL announces its version on stdout, which must
        # carry ONE JSON line, so fd 1 points at /dev/null meanwhile
        _quiet_call(lambda: (dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank)), dist.barrier()))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one codec context (= its own CUDA streams and scratch) per level: the three levels of a step are
    # independent jobs, so they are issued concurrently from host threads, see PIPELINES (ctypes drops the GIL)
    cds = {lv: xpng_b200.Codec(local_rank) for lv in LEVELS}
    cd = cds[1]
    stream = torch.cuda.ExternalStream(cd.stream, device=dev)
    lib = xpng_b200.lib()
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=len(PIPELINES))

    frame = synth.rgb(H, W, 1 + rank)
    npx = W * H
    descs, total = xpng_b200.Codec.layout([frame.shape])
    cap = int(lib.xpngb_encode_bound(descs, 1))
    # device-resident buffers
    d_px = torch.from_numpy(frame.reshape(-1)).to(dev)
    d_px = torch.cat([d_px, torch.zeros(64, dtype=torch.uint8, device=dev)])
    d_files = {lv: torch.zeros(cap + 64, dtype=torch.uint8, device=dev) for lv in LEVELS}
    d_back = {lv: torch.zeros(total + 64, dtype=torch.uint8, device=dev) for lv in LEVELS}
    # pinned host buffers for the e2e leg
    h_px = torch.from_numpy(frame.reshape(-1).copy()).pin_memory()
    h_files = {lv: torch.zeros(cap + 64, dtype=torch.uint8).pin_memory() for lv in LEVELS}
    h_back = {lv: torch.zeros(total + 64, dtype=torch.uint8).pin_memory() for lv in LEVELS}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    sizes = {}
    launches = [0]

    def enc_one(lv, px, files, on_dev):
        d = xpng_b200.Codec.layout([frame.shape])[0]
        offs, sz = cds[lv].encode_raw(lv, d, 1, px.data_ptr(), total, on_dev, files[lv].data_ptr(), cap, on_dev)
        sizes[lv] = (int(offs[0]), int(sz[0]))
        return cds[lv].last_launches

    def dec_one(lv, files, back, on_dev):
        d = xpng_b200.Codec.layout([frame.shape])[0]
        d[0].w = d[0].h = 0
        off = (C.c_uint64 * 1)(sizes[lv][0]); sz = (C.c_uint64 * 1)(sizes[lv][1])
        cds[lv].decode_raw(d, 1, files[lv].data_ptr(), cap, on_dev, off, sz, back[lv].data_ptr(), total, on_dev)
        return cds[lv].last_launches

    def step(px, files, back, on_dev):
        # per level: encode, then decode of that level's file; the pipelines of PIPELINES are in flight together
        launches[0] += sum(pool.map(lambda lvs: sum(enc_one(lv, px, files, on_dev) + dec_one(lv, files, back, on_dev) for lv in lvs), PIPELINES))

    def timed(px, files, back, on_dev, steps, warmup):
        for _ in range(warmup):
            step(px, files, back, on_dev)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        launches[0] = 0
        ms = 0.0
        for _ in range(steps):
            flush.fill_(1)                      # L2 flush between timed iterations (outside the events)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step(px, files, back, on_dev)
            e1.record(stream)
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # parity gate before any timing counts: bytes and pixels against the oracle
    step(d_px, d_files, d_back, 1)
    for lv in LEVELS:
        o, s = sizes[lv]
        got = d_files[lv][o:o + s].cpu().numpy().tobytes()
        assert got == po.encode(lv, frame), f"level {lv}: bytes differ from the oracle"
    for lv in LEVELS:
        assert np.array_equal(d_back[lv][:frame.size].cpu().numpy(), frame.reshape(-1)), f"level {lv}: decoded pixels differ"
    xpng_bytes = {lv: sizes[lv][1] for lv in LEVELS}

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev = timed(d_px, d_files, d_back, 1, args.steps, args.warmup)
    n_launch = launches[0]
    clocks = sampler.stop()
    ms_e2e = timed(h_px, h_files, h_back, 0, args.steps, max(1, args.warmup))
    for lv in LEVELS:
        assert np.array_equal(h_back[lv][:frame.size].numpy(), frame.reshape(-1))

    pix_per_step = 6 * npx * world
    value = pix_per_step / 1e6 / (ms_dev / args.steps / 1e3)
    e2e = pix_per_step / 1e6 / (ms_e2e / args.steps / 1e3)
    raw = frame.size
    h2d = 3 * raw + sum(xpng_bytes.values())      # encodes upload the pixels, decodes upload the files
    d2h = sum(xpng_bytes.values()) + 3 * raw

    if rank != 0:
        return
    # ---- per-op breakdown and the dominant kernel (serialised profiling pass, outside the timed region)
    breakdown = {}
    cd.profile(True)
    per_kernel = {}
    for lv in LEVELS:
        for what in ("enc", "dec"):
            cd.profile(True)
            t0 = time.perf_counter()
            d = xpng_b200.Codec.layout([frame.shape])[0]
            if what == "enc":
                cd.encode_raw(lv, d, 1, d_px.data_ptr(), total, 1, d_files[lv].data_ptr(), cap, 1)
            else:
                d[0].w = d[0].h = 0
                off = (C.c_uint64 * 1)(sizes[lv][0]); sz = (C.c_uint64 * 1)(sizes[lv][1])
                cd.decode_raw(d, 1, d_files[lv].data_ptr(), cap, 1, off, sz, d_back[lv].data_ptr(), total, 1)
            rep = cd.profile_report()
            for k, (ms, cnt) in rep.items():
                per_kernel[f"L{lv}.{what}.{k}"] = (ms / cnt, lv, what)
            breakdown[f"{what}{lv}_kernel_ms"] = round(sum(ms for ms, _ in rep.values()), 3)
    cd.profile(False)
    for lv in LEVELS:      # un-profiled per-call times (device resident)
        for what in ("enc", "dec"):
            best = 1e9
            for _ in range(3):
                d = xpng_b200.Codec.layout([frame.shape])[0]
                torch.cuda.synchronize(); t0 = time.perf_counter()
                if what == "enc":
                    cd.encode_raw(lv, d, 1, d_px.data_ptr(), total, 1, d_files[lv].data_ptr(), cap, 1)
                else:
                    d[0].w = d[0].h = 0
                    off = (C.c_uint64 * 1)(sizes[lv][0]); sz = (C.c_uint64 * 1)(sizes[lv][1])
                    cd.decode_raw(d, 1, d_files[lv].data_ptr(), cap, 1, off, sz, d_back[lv].data_ptr(), total, 1)
                best = min(best, time.perf_counter() - t0)
            breakdown[f"{what}{lv}_MPix_s"] = round(npx / 1e6 / best, 1)
    top = max(per_kernel.items(), key=lambda kv: kv[1][0])
    top_name, (top_ms, top_lv, top_what) = top
    peak, peak_src = peaks()
    alg_bytes = raw + xpng_bytes[top_lv] if top_lv != 7 else 2 * raw
    achieved = alg_bytes / 1e9 / (top_ms / 1e3)
    # DRAM traffic of that kernel from the committed `ncu --set full` capture (profiles/r01_traffic.json), per launch
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        key = {"k_rans_v1_pair_16": "k_rans_v1_pair<16>", "k_rans_v2_pair_16": "k_rans_v2_pair<16>", "k_dec_walk_smem<0>": "k_dec_walk_smem<0>",
               "k_dec_rans_v1_lat_values16": "k_dec_rans_v1_lat", "k_dec_rans_v2_lat": "k_dec_rans_v2_lat"}.get(top_name.split(".")[-1])
        for name, rec in tj.items():
            if key and key in name:
                traffic = int(rec["dram_read_bytes"] + rec["dram_write_bytes"]); traffic_src = "profiles/" + rec["report"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": top_name, "kernel_ms": round(top_ms, 4), "achieved": round(achieved, 2), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 5), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "dominant kernel is a serial chain fixed by the bit stream (2 rANS states per entropy block): one 4K frame is "
                        "latency-bound, 48 warps on 148 SMs; see DESIGN.md section 3 and 6, profiles/r01_ncu_summaries.md"}
    # ---- whole-step roofline view: algorithmic bytes of all 6 calls over the step time
    step_bytes = sum((raw + xpng_bytes[lv]) if lv != 7 else 2 * raw for lv in LEVELS) * 2
    breakdown["step_algorithmic_GB_s"] = round(step_bytes / 1e9 / (ms_dev / args.steps / 1e3), 2)

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        times, kind, cores, sample = reference_steps(frame, 3, 1)
        per = sum(times) / len(times)
        cpu = {"value": round(6 * npx / 1e6 / per, 2), "unit": "MPix/s", "cores": cores, "kind": kind, "sample": sample + "; 3 steps after 1 warm-up"}

    line = {"metric": METRIC, "value": round(value, 2), "unit": "MPix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_dev / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "configs[1]: one 3840x2160 RGB synthetic frame per GPU (seed 1+rank), levels -1/-2/-7, encode+decode",
                       "frames_per_gpu": 1, "tiles_per_frame": 45, "l2": "flushed between timed steps (256 MiB fill)",
                       "concurrency": "the 3 levels of a step run as 2 concurrent encode->decode pipelines: level 2 | level 1 then level 7 (one codec context and CUDA stream set per level)",
                       "parallelism": f"frames sharded over {world} GPU(s), no collective"},
            "e2e": {"value": round(e2e, 2), "unit": "MPix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(ms_e2e / args.steps, 4)},
            "gpu_launches": n_launch, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "xpng_bytes": xpng_bytes, "breakdown": breakdown}
    print(json.dumps(line), flush=True)
    if dist is not None:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    run_ours(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
