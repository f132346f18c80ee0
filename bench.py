#!/usr/bin/env python
"""bench.py -- throughput of the Dr.Jit data-parallel primitive path on B200.

One "step" = one pass of the primitive suite over synthetic fmix32 inputs resident in HBM
(generator: ext/drjit-core/tests/reductions.cpp:5-13; sizes: BASELINE.json configs):

    sum f32 2^28 | block_reduce(Add,256) f32 2^28 | dot f32 2^28 | exclusive prefix_sum u32 2^30 |
    compress 2^30 (50 % dense) | block_mkperm 2^26 keys x 4096 buckets | scatter_add f32 2^28 -> 2^20 bins

value = algorithmic bytes of the whole suite / device time of one step (GB/s), inputs larger
than L2 (no flush needed). With --gpus N every rank holds one shard of the sizes above (the
global arrays are N times larger: "weak" scaling, rank r owns the contiguous index range
[r*n, (r+1)*n)); NCCL is used only for the combine messages.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scale S]

--impl reference times the reference's own CPU implementation of the path (the unmodified
drjit-core LLVM-backend primitives built into oracle/_ref) on the host cores.
--scale S shrinks every array by 2^S (debugging only; the JSON says so).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "achieved HBM GB/s & % peak: reduce/scan/compress/mkperm @2^28, 1/2/4/8 GPU"

# name -> (log2 elements, algorithmic bytes per element)   [SURVEY.md section 8d / BASELINE.md 2c]
SUITE = [
    ("sum_f32", 28, 4.0),
    ("block_reduce256_f32", 28, 4.0 + 4.0 / 256),
    ("dot_f32", 28, 8.0),
    ("prefix_sum_u32", 30, 8.0),
    ("compress_u8", 30, 3.0),          # 1 + 4 * density, density = 0.5
    ("mkperm_4096", 26, 12.0),
    ("scatter_add_f32", 28, 8.0),
]
BINS_LOG2 = 20
BUCKETS = 4096


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
#  clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls nvidia-smi every 100 ms from before the warm-up until after the timed region. The
    median SM clock is taken over the samples that fall inside the timed region; when that
    region is shorter than a few polling periods, over warm-up + timed steps (same load), and
    `window` says which."""

    def __init__(self, index):
        self.index = index
        self.samples = []       # (monotonic time, sm MHz, max MHz)
        self.reasons = []       # (monotonic time, name)
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def wait_first(self, timeout=10.0):
        """nvidia-smi needs a second or so to start: block until it delivers"""
        t0 = time.monotonic()
        while self.proc and not self.samples and time.monotonic() - t0 < timeout:
            time.sleep(0.05)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            now = time.monotonic()
            try:
                self.samples.append((now, float(parts[0]), float(parts[1])))
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.append((now, n))
            except (ValueError, IndexError):
                pass

    def stop(self, load_start=None, timed_start=None, timed_end=None):
        if self.proc:
            self.proc.terminate()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "window": "nvidia-smi gave no samples"}
        window = "timed steps"
        sel = [s for s in self.samples if timed_start is not None and timed_start <= s[0] <= timed_end]
        lo, hi = timed_start, timed_end
        if len(sel) < 3 and load_start is not None:
            window = "warm-up + timed steps (timed region shorter than 3 polling periods)"
            sel = [s for s in self.samples if load_start <= s[0] <= timed_end]
            lo = load_start
        if not sel:
            window, sel, lo, hi = "whole run", self.samples, self.samples[0][0], self.samples[-1][0]
        sm = sorted(s[1] for s in sel)
        reasons = sorted({n for t, n in self.reasons if lo <= t <= hi})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": sel[0][2], "reasons": reasons,
                "samples": len(sel), "window": window}


def measured_traffic(kernel, elements):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum); None when the capture
    was taken at another size."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            e = json.load(f)[kernel]
        return float(e["dram_bytes_per_launch"]) if int(e["elements"]) == int(elements) else None
    except (OSError, KeyError, ValueError):
        return None


# ------------------------------------------------------------------------------------------
#  reference arm: the unmodified reference's CPU primitives (oracle/_ref), all host threads
# ------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_log2):
    """Times the suite on a bounded sample (2^sample_log2 elements per primitive, 2^(sample-2)
    for mkperm) with the reference LLVM backend. Returns (GB/s, per-primitive dict, cores)."""
    import numpy as np
    from oracle import capi, ref

    cores = os.cpu_count() or 1
    L = ref.lib(cuda=False, llvm=True)
    L.ref_llvm_set_thread_count(cores)
    n = 1 << sample_log2
    nk = 1 << max(10, sample_log2 - 2)
    x = capi.unit_f32(n); y = capi.unit_f32(n, xor=0x9E3779B9)
    u = capi.fmix32(n); m = capi.mask_u8(n, 128); keys = capi.fmix32(nk, mask=BUCKETS - 1)

    import ctypes
    vp = ctypes.c_void_p
    P = lambda a: a.ctypes.data_as(vp)  # noqa: E731
    out_f = np.zeros(max(1, n // 256), np.float32); out_u = np.empty(n, np.uint32)
    idx = np.empty(n, np.uint32); perm = np.empty(nk, np.uint32); offs = np.zeros(4 * BUCKETS + 1, np.uint32)
    dot_out = np.zeros(1, np.float32)
    VT_F32, VT_U32, ADD, LLVM = 14, 8, 1, 2

    prims = {
        "sum_f32": (n * 4.0, lambda: (L.ref_block_reduce(LLVM, VT_F32, ADD, n, n, P(x), P(out_f)), L.ref_sync())),
        "block_reduce256_f32": (n * (4.0 + 4.0 / 256), lambda: (L.ref_block_reduce(LLVM, VT_F32, ADD, n, 256, P(x), P(out_f)), L.ref_sync())),
        "dot_f32": (n * 8.0, lambda: L.ref_reduce_dot(LLVM, VT_F32, P(x), P(y), n, P(dot_out))),
        "prefix_sum_u32": (n * 8.0, lambda: (L.ref_block_prefix_reduce(LLVM, VT_U32, ADD, n, n, 1, 0, P(u), P(out_u)), L.ref_sync())),
        "compress_u8": (n * 3.0, lambda: L.ref_compress(LLVM, P(m), n, P(idx))),
        "mkperm_4096": (nk * 12.0, lambda: (L.ref_block_mkperm(LLVM, P(keys), nk, nk, BUCKETS, P(perm), P(offs)), L.ref_sync())),
        # scatter_add: the reference CPU path needs its LLVM JIT (ReduceMode::Expand), which the
        # stub libLLVM cannot provide -> reported as n/a and left out of the CPU aggregate.
    }
    times = {k: [] for k in prims}
    for it in range(warmup + steps):
        for name, (_, fn) in prims.items():
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times[name].append(dt)
    per = {}
    tot_bytes = tot_time = 0.0
    for name, (nbytes, _) in prims.items():
        t = sum(times[name]) / len(times[name])
        per[name] = {"GBps": nbytes / t / 1e9, "ms": t * 1e3}
        tot_bytes += nbytes; tot_time += t
    per["scatter_add_f32"] = "n/a (reference CPU scatter needs the LLVM JIT)"
    return tot_bytes / tot_time / 1e9, per, cores, tot_time


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_log2 = 26
    value, per, cores, step_s = cpu_reference_run(args.steps, max(1, min(args.warmup, 2)), sample_log2)
    sample = (f"suite without scatter_add at 2^{sample_log2} elements per primitive (mkperm 2^{sample_log2 - 2} keys), "
              f"host arrays, unmodified reference LLVM-backend primitives, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32/u32/u8", "data": "synthetic (fmix32)",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "primitives": per, "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "primitive suite: sum/block_reduce(256)/dot f32 2^28, exclusive prefix_sum u32 2^30, "
                        "compress u8 2^30 (50%), block_mkperm 2^26 x 4096 buckets, scatter_add f32 2^28 -> 2^20 bins",
            "sharding": f"{args.gpus} rank(s), each owning one contiguous shard of the sizes above (global arrays "
                        f"are {args.gpus}x larger; compress indices are global mod 2^32); NCCL only for combine messages"
                        + ("; prefix_sum in shard-offset form (local scan + per-shard offset, like the compress / mkperm "
                           "offsets), materialised form reported alongside" if args.gpus > 1 else ""),
            "l2": "every input array > 126 MB L2 (no flush needed)",
            "scale_shift": args.scale}


# ------------------------------------------------------------------------------------------
#  GPU arm
# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200,
                    help="timed steps (default 200: ~1 s, long enough for nvidia-smi to sample the clocks inside the timed region)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=0, help="shrink every array by 2^SCALE (debug)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import drjit_b200 as dr
    from drjit_b200 import ReduceOp, VarType, ops
    from drjit_b200 import dist as ddist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sh = ddist.Sharded(rank=rank, world=world, group=dist.group.WORLD if world > 1 else None)

    peak, peak_src = measured_peaks()
    S = args.scale
    size = {name: 1 << (lg - S) for name, lg, _ in SUITE}
    bpe = {name: b for name, _, b in SUITE}
    bins = 1 << max(4, BINS_LOG2 - S)

    # ---- shard-resident synthetic inputs (generated on the device) -------------------------
    def shard(n):
        """Weak scaling: rank r owns elements [r*n, (r+1)*n) of a global array of world*n entries."""
        return rank * n, (rank + 1) * n

    lo_f, hi_f = shard(size["sum_f32"])
    x = torch.empty(hi_f - lo_f, dtype=torch.float32, device=dev); ops.fill_fmix32(x, 1, start=lo_f)
    y = torch.empty(hi_f - lo_f, dtype=torch.float32, device=dev); ops.fill_fmix32(y, 1, start=lo_f, xor=0x9E3779B9)
    lo_u, hi_u = shard(size["prefix_sum_u32"])
    u = torch.empty(hi_u - lo_u, dtype=torch.int32, device=dev); ops.fill_fmix32(u, 0, start=lo_u)
    u_out = torch.empty_like(u)
    lo_m, hi_m = shard(size["compress_u8"])
    mask = torch.empty(hi_m - lo_m, dtype=torch.uint8, device=dev); ops.fill_fmix32(mask, 2, start=lo_m, and_=128)
    c_out = torch.empty(hi_m - lo_m, dtype=torch.int32, device=dev)
    lo_k, hi_k = shard(size["mkperm_4096"])
    keys = torch.empty(hi_k - lo_k, dtype=torch.int32, device=dev); ops.fill_fmix32(keys, 0, start=lo_k, and_=BUCKETS - 1)
    perm = torch.empty_like(keys)
    lo_s, hi_s = shard(size["scatter_add_f32"])
    sidx = torch.empty(hi_s - lo_s, dtype=torch.int32, device=dev); ops.fill_fmix32(sidx, 0, start=lo_s, xor=0x85EBCA6B, and_=bins - 1)
    sval = x if (lo_s, hi_s) == (lo_f, hi_f) else torch.empty(hi_s - lo_s, dtype=torch.float32, device=dev)
    if sval is not x:
        ops.fill_fmix32(sval, 1, start=lo_s)
    bins_t = torch.zeros(bins, dtype=torch.float32, device=dev)
    br_out = torch.empty((x.numel() + 255) // 256, dtype=torch.float32, device=dev)

    results = {}

    def p_sum():
        results["sum"] = sh.reduce(ReduceOp.Add, x)

    def p_block_reduce():
        results["br"] = ops.block_reduce(ReduceOp.Add, x, 256, out=br_out)   # block-aligned shards: no exchange

    def p_dot():
        results["dot"] = sh.dot(x, y)

    def p_prefix():
        # N > 1: shard-offset form (local scan + per-shard offset, the representation the compress /
        # mkperm offsets use as well; dist.py). The materialised form is timed separately below.
        if world > 1:
            results["scan"] = sh.prefix_reduce_offsets(ReduceOp.Add, u, vt=VarType.UInt32, out=u_out)
        else:
            results["scan"] = sh.prefix_sum(u, vt=VarType.UInt32, out=u_out)

    def p_compress():
        results["count"] = sh.compress(mask, lo_m & 0xFFFFFFFF, out=c_out)

    def p_mkperm():
        results["mkperm"] = sh.mkperm(keys, BUCKETS, lo_k & 0xFFFFFFFF, perm=perm)

    def p_scatter():
        bins_t.zero_()
        results["bins"] = sh.scatter_add(bins_t, sval, sidx)

    prims = [("sum_f32", p_sum), ("block_reduce256_f32", p_block_reduce), ("dot_f32", p_dot),
             ("prefix_sum_u32", p_prefix), ("compress_u8", p_compress), ("mkperm_4096", p_mkperm),
             ("scatter_add_f32", p_scatter)]
    total_bytes = world * sum(size[n] * bpe[n] for n, _ in prims)   # all ranks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(events=None):
        for i, (name, fn) in enumerate(prims):
            if events is not None:
                events[i][0].record()
            fn()
            if events is not None:
                events[i][1].record()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    barrier()
    t_load = time.monotonic()
    # W warm-up steps, padded (same count on every rank) so that the clocks are sampled under
    # this load for >= 1 s before the timed region starts
    for _ in range(args.warmup):
        one_step()
    barrier()
    elapsed = time.monotonic() - t_load
    extra = 0 if elapsed >= 1.0 else min(2000, int((1.0 - elapsed) / max(elapsed / args.warmup, 1e-4)) + 1)
    if world > 1:
        t = torch.tensor([extra], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        extra = int(t.item())
    for _ in range(extra):
        one_step()
    barrier()
    dr.launch_count(reset=True)
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in prims]
          for _ in range(args.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_timed0 = time.monotonic()
    start.record()
    for k in range(args.steps):
        one_step(ev[k])
    end.record()
    barrier()
    t_timed1 = time.monotonic()
    launches = dr.launch_count()
    clocks = sampler.stop(t_load, t_timed0, t_timed1) if rank == 0 else None

    total_ms = start.elapsed_time(end)
    per_ms = [sum(ev[k][i][0].elapsed_time(ev[k][i][1]) for k in range(args.steps)) / args.steps
              for i in range(len(prims))]
    t = torch.tensor([total_ms] + per_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t[0]); per_ms = [float(v) for v in t[1:]]
    ms_per_step = total_ms / args.steps
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    primitives = {}
    for (name, _), ms in zip(prims, per_ms):
        gbs = world * size[name] * bpe[name] / (ms * 1e-3) / 1e9
        primitives[name] = {"elements": world * size[name], "ms": round(ms, 4), "GBps": round(gbs, 1),
                            "Gelem_per_s": round(world * size[name] / (ms * 1e-3) / 1e9, 2),
                            "frac_of_peak_per_gpu": round(gbs / (peak * world), 4)}

    # ---- informational (NOT part of `value`): the sharded scan in materialised form (every element
    # carries the global value: one more read pass over the shard, 12 B/element) next to the
    # shard-offset form timed above (8 B/element)
    if world > 1:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sh.prefix_sum(u, vt=VarType.UInt32, out=u_out)
        barrier()
        a.record()
        for _ in range(args.steps):
            sh.prefix_sum(u, vt=VarType.UInt32, out=u_out)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        gbs = world * size["prefix_sum_u32"] * 8.0 / (ms * 1e-3) / 1e9
        primitives["prefix_sum_u32"]["form"] = "shard-offset (local scan + per-shard offset)"
        primitives["prefix_sum_u32"]["materialised_form"] = {
            "ms": round(ms, 4), "GBps": round(gbs, 1), "frac_of_peak_per_gpu": round(gbs / (peak * world), 4),
            "note": "global value in every element (extra read pass over the shard); informational, not in `value`"}

    # ---- end-to-end: host buffers through the public API, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, dr, ops, sh, dev, world, rank,
                      dict(x=x, y=y, u=u, mask=mask, keys=keys, sidx=sidx, sval=sval),
                      dict(u_out=u_out, c_out=c_out, perm=perm, bins_t=bins_t, br_out=br_out),
                      prims, results, total_bytes)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # dominant kernel for the roofline object: the single-pass scan (largest share of the step)
    scan = primitives["prefix_sum_u32"]
    roofline = {"bound": "hbm", "kernel": "prefix_reduce_kernel<u32,Add> (exclusive prefix_sum, 2^30 elements per rank)",
                "achieved": scan["GBps"] / world, "peak": peak, "unit": "GB/s",
                "frac": round(scan["GBps"] / world / peak, 4),
                "traffic": measured_traffic("prefix_reduce_kernel", size["prefix_sum_u32"]), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": size["prefix_sum_u32"] * 8.0}

    cpu_baseline = None
    if not args.no_cpu:
        try:
            v, per, cores, _ = cpu_reference_run(2, 1, 24)
            cpu_baseline = {"value": round(v, 2), "unit": "GB/s", "cores": cores, "kind": "reference",
                            "sample": "suite without scatter_add at 2^24 elements per primitive (mkperm 2^22 keys), "
                                      "unmodified reference LLVM-backend CPU primitives (oracle/_ref), all host threads",
                            "primitives": {k: (round(p["GBps"], 2) if isinstance(p, dict) else p) for k, p in per.items()}}
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference",
                            "sample": f"unavailable: {e}"}

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32/u32/u8", "data": "synthetic (fmix32)",
        "config": workload_config(args), "frac_of_hbm_peak": round(value / (peak * world), 4),
        "primitives": primitives, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, torch, dist, dr, ops, sh, dev, world, rank, inputs, outputs, prims, results, total_bytes):
    """Same suite, but every step first copies the step's inputs host->device from pinned
    memory and afterwards reads every primitive's result back to the host."""
    from drjit_b200 import ReduceOp, VarType
    host_in = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in inputs.items() if k != "sval" or v is not inputs["x"]}
    for k, h in host_in.items():
        h.copy_(inputs[k])
    host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in outputs.items()}
    h2d = sum(h.numel() * h.element_size() for h in host_in.values())
    steps = max(2, min(args.steps, 3))

    # Pipelined step: uploads, kernels and downloads run on three streams (PCIe is full duplex and
    # the GPU has separate copy engines per direction), chained by events per primitive. The
    # primitives with the largest results go first so that their downloads overlap the remaining
    # uploads. Every input is still copied host->device and every result device->host in every step.
    fn = dict(prims)
    plan = [  # (primitive, inputs it needs uploaded, outputs to download)
        ("compress_u8", ["mask"], ["c_out"]),          # 1 GB up, 2 GB down: its download overlaps the next upload
        ("prefix_sum_u32", ["u"], ["u_out"]),         # 4 GB up, 4 GB down
        ("sum_f32", ["x"], []),
        ("block_reduce256_f32", [], ["br_out"]),
        ("dot_f32", ["y"], []),
        ("mkperm_4096", ["keys"], ["perm"]),
        ("scatter_add_f32", ["sidx"] + (["sval"] if "sval" in host_in else []), ["bins_t"]),
    ]
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)

    # The scan is fed in SCAN_CHUNKS pieces through the carry form of the primitive
    # (ops.prefix_reduce_carry: running value in, running total out), so that the upload of piece
    # k+1, the scan of piece k and the download of piece k-1 overlap: its 4 GB up and 4 GB down then
    # travel at the same time instead of one after the other.
    SCAN_CHUNKS = 8
    u_dev, u_res = inputs["u"], outputs["u_out"]
    n_u = u_dev.numel()
    cs = (n_u // SCAN_CHUNKS + 3) // 4 * 4 if n_u >= 64 * SCAN_CHUNKS else n_u
    bounds = [(lo, min(lo + cs, n_u)) for lo in range(0, n_u, cs)]
    carry = [torch.zeros(1, dtype=u_dev.dtype, device=dev), torch.zeros(1, dtype=u_dev.dtype, device=dev)]

    def scan_chunked(up_u):
        moved = 0
        for c, (lo, hi) in enumerate(bounds):
            main.wait_event(up_u[c])
            ops.prefix_reduce_carry(ReduceOp.Add, u_dev[lo:hi], True, False, carry_in=carry[c & 1] if c else None,
                                    total_out=carry[(c + 1) & 1], vt=VarType.UInt32, out=u_res[lo:hi])
            done = torch.cuda.Event(); done.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                host_out["u_out"][lo:hi].copy_(u_res[lo:hi], non_blocking=True)
            moved += (hi - lo) * u_res.element_size()
        if world > 1:       # shard-offset form: exclusive scan of the gathered shard totals
            totals = sh._all_gather(carry[len(bounds) & 1])
            offs = ops.block_prefix_reduce(ReduceOp.Add, totals, totals.numel(), True, False, vt=VarType.UInt32)
            results["scan"] = (u_res, offs[rank:rank + 1])
        else:
            results["scan"] = u_res
        return moved

    def step():
        main_ready = torch.cuda.Event(); main_ready.record(main)
        up = {}
        with torch.cuda.stream(s_in):
            s_in.wait_event(main_ready)            # previous step's kernels are done with the inputs
            for _, ins, _ in plan:
                for k in ins:
                    if k == "u":
                        up[k] = []
                        for lo, hi in bounds:
                            u_dev[lo:hi].copy_(host_in[k][lo:hi], non_blocking=True)
                            e = torch.cuda.Event(); e.record(s_in); up[k].append(e)
                        continue
                    inputs[k].copy_(host_in[k], non_blocking=True)
                    up[k] = torch.cuda.Event(); up[k].record(s_in)
        d2h = 0
        for name, ins, outs in plan:
            if name == "prefix_sum_u32":
                d2h += scan_chunked(up["u"])
                continue
            for k in ins:
                main.wait_event(up[k])
            fn[name]()
            done = torch.cuda.Event(); done.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                for k in outs:
                    v = outputs[k]
                    if k == "c_out":
                        count = results["count"][1][rank]
                        host_out[k][:count].copy_(v[:count], non_blocking=True); d2h += count * 4
                    else:
                        host_out[k].copy_(v, non_blocking=True); d2h += v.numel() * v.element_size()
        for k in ("sum", "dot"):
            results[k].cpu(); d2h += 4
        if isinstance(results["scan"], tuple):      # shard-offset form: the offset travels with the scan
            results["scan"][1].cpu(); d2h += 4
        s_out.synchronize()
        return d2h

    step()
    torch.cuda.synchronize()
    # the pieces must add up to the one-shot scan of the shard (checked once, outside the timed region)
    whole = ops.prefix_reduce_carry(ReduceOp.Add, u_dev, True, False, vt=VarType.UInt32)
    if not torch.equal(whole, u_res) or not torch.equal(host_out["u_out"], whole.cpu()):
        raise SystemExit("bench.py: chunked scan differs from the one-shot scan")
    del whole
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(steps):
        d2h = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / steps
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
    return {"value": round(total_bytes / dt / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": int(h2d) * world,
            "d2h_bytes_per_step": int(d2h) * world, "ms_per_step": round(dt * 1e3, 3), "steps": steps,
            "note": "per rank: pinned host inputs -> device, suite through the public API, every result "
                    "(scalars, block sums, scan, index list, permutation, bins) -> pinned host; uploads, "
                    "kernels and downloads pipelined on three streams, the scan fed in 8 pieces through its carry form"}


if __name__ == "__main__":
    main()
