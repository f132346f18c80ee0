#!/usr/bin/env python
"""bench.py -- throughput of the Dr.Jit data-parallel primitive path on B200.

One "step" = one pass of the primitive suite over synthetic fmix32 inputs resident in HBM
(generator: ext/drjit-core/tests/reductions.cpp:5-13; sizes: BASELINE.json configs):

    sum f32 2^28 | block_reduce(Add,256) f32 2^28 | dot f32 2^28 | exclusive prefix_sum u32 2^30 |
    compress 2^30 (50 % dense) | block_mkperm 2^26 keys x 4096 buckets | scatter_add f32 2^28 -> 2^20 bins

value = algorithmic bytes of the whole suite / device time of one step (GB/s).

--gpus N (one rank per GPU, torchrun): STRONG scaling -- the global arrays keep the sizes above
and rank r owns the contiguous index range [r*n/N, (r+1)*n/N) (north_star: "2^30-element mask
sharded across 1/2/4/8"). The combine step of every primitive (a scalar, N counts, a bucket
histogram, the 4 MB bin array) runs INSIDE the primitive's kernel over peer-mapped NVLink windows
(drjit_b200/csrc/comm.cuh); the same suite over NCCL collectives is timed beside it
(`nccl_path_ms`). After the timed region one untimed pass is verified against torch-computed
invariants on every rank ("verified"). --weak gives every rank a full-size shard instead; a short
weak-scaling run is also appended to the strong-scaling line (`weak_scaling`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--weak] [--scale S]

--impl reference times the reference's own CPU implementation of the path (the unmodified
drjit-core LLVM-backend primitives built into oracle/_ref) on the host cores at the same sizes.
--scale S shrinks every array by 2^S (debugging only; the JSON says so).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "achieved HBM GB/s & % peak: reduce/scan/compress/mkperm @2^28, 1/2/4/8 GPU"

# name, log2 elements, algorithmic bytes per element [SURVEY.md section 8d / BASELINE.md 2c], kernel
SUITE = [
    ("sum_f32", 28, 4.0, "block_reduce_chunk_kernel<f32,Add>"),
    ("block_reduce256_f32", 28, 4.0 + 4.0 / 256, "block_reduce_group_kernel<f32,Add,16 lanes>"),
    ("dot_f32", 28, 8.0, "block_reduce_chunk_kernel<f32,Add,dot>"),
    ("prefix_sum_u32", 30, 8.0, "prefix_reduce_kernel<u32,Add>"),
    ("compress_u8", 30, 3.0, "compress_kernel<8,1,3,kCopyLsuPairs>"),          # 1 + 4 * density, density = 0.5
    ("mkperm_4096", 26, 12.0, "mkperm_tile_hist_kernel + column/bucket scan kernels + mkperm_tile_scatter_kernel<1024,48>"),
    ("scatter_add_f32", 28, 8.0, "scatter_reduce_kernel<f32,Add>"),
]
BINS_LOG2 = 20
BUCKETS = 4096
L2_BYTES = 126 << 20


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


def measured_traffic(primitive, elements):
    """DRAM bytes per call from the committed `ncu --set full` captures (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum over the primitive's kernels); None when no
    capture exists at this size."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            e = json.load(f)[primitive]
        return float(e["dram_bytes_per_launch"]) if int(e["elements"]) == int(elements) else None
    except (OSError, KeyError, ValueError):
        return None


def workload_text():
    return ("primitive suite: sum/block_reduce(256)/dot f32 2^28, exclusive prefix_sum u32 2^30, "
            "compress u8 2^30 (50%), block_mkperm 2^26 x 4096 buckets, scatter_add f32 2^28 -> 2^20 bins")


def workload_config(args, weak=False, distinct=False):
    n = args.gpus
    if weak:
        sharding = (f"weak: {n} rank(s), each owning a full-size shard (global arrays {n}x larger; "
                    "compress indices are global mod 2^32)")
    else:
        sharding = (f"strong: global sizes as stated, rank r of {n} owns the contiguous range [r*n/{n}, (r+1)*n/{n})")
    if n > 1:
        sharding += ("; combine steps fused into the primitives' kernels over peer-mapped NVLink windows (no NCCL on "
                     "the data path); prefix_sum in shard-offset form (local scan + per-shard offset, like the "
                     "compress / mkperm offsets), materialised form reported alongside")
    l2 = ("no flush: every input array a primitive reads is larger than the 126 MB L2, or -- when the per-rank shards "
          "approach the L2 size (N >= 4) -- every primitive reads its own input arrays and one step touches > 2 GB per "
          "rank between re-uses, so no input is L2-resident when its primitive starts")
    return {"workload": workload_text(), "sharding": sharding, "l2": l2, "scale_shift": args.scale}


# ------------------------------------------------------------------------------------------
#  reference arm: the unmodified reference's CPU primitives (oracle/_ref), all host threads
# ------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, scale):
    """Times the suite at the BASELINE sizes (shrunk by 2^scale) with the reference LLVM-backend
    primitives on host arrays, all host threads. scatter_add: the reference JIT-compiles that
    loop, which the stub libLLVM of oracle/ref_build cannot do -> oracle.c's restatement of the
    reference's ReduceMode::Expand scheme (one private copy of the bins per worker + fold) with the
    same thread count. Returns (GB/s, per-primitive dict, cores, seconds per step)."""
    import ctypes

    import numpy as np
    from oracle import capi, ref

    cores = os.cpu_count() or 1
    L = ref.lib(cuda=False, llvm=True)
    L.ref_llvm_set_thread_count(cores)
    size = {name: 1 << max(10, lg - scale) for name, lg, _, _ in SUITE}
    bpe = {name: b for name, _, b, _ in SUITE}
    bins = 1 << max(4, BINS_LOG2 - scale)
    n, ns, nk = size["sum_f32"], size["prefix_sum_u32"], size["mkperm_4096"]
    x = capi.unit_f32(n); y = capi.unit_f32(n, xor=0x9E3779B9)
    u = capi.fmix32(ns); m = capi.mask_u8(ns, 128); keys = capi.fmix32(nk, mask=BUCKETS - 1)
    sidx = capi.fmix32(n, xor=0x85EBCA6B, mask=bins - 1)

    vp = ctypes.c_void_p
    P = lambda a: a.ctypes.data_as(vp)  # noqa: E731
    out_f = np.zeros(max(1, n // 256), np.float32); out_u = np.empty(ns, np.uint32)
    idx = np.empty(ns, np.uint32); perm = np.empty(nk, np.uint32); offs = np.zeros(4 * BUCKETS + 1, np.uint32)
    dot_out = np.zeros(1, np.float32)
    bins_a = np.zeros(bins, np.float32); scratch = np.empty(cores * bins, np.float32)
    VT_F32, VT_U32, ADD, LLVM = 14, 8, 1, 2

    def scatter():
        bins_a[:] = 0
        capi.scatter_add_expand_f32(bins_a, x, sidx, cores, scratch)

    prims = {
        "sum_f32": lambda: (L.ref_block_reduce(LLVM, VT_F32, ADD, n, n, P(x), P(out_f)), L.ref_sync()),
        "block_reduce256_f32": lambda: (L.ref_block_reduce(LLVM, VT_F32, ADD, n, 256, P(x), P(out_f)), L.ref_sync()),
        "dot_f32": lambda: L.ref_reduce_dot(LLVM, VT_F32, P(x), P(y), n, P(dot_out)),
        "prefix_sum_u32": lambda: (L.ref_block_prefix_reduce(LLVM, VT_U32, ADD, ns, ns, 1, 0, P(u), P(out_u)), L.ref_sync()),
        "compress_u8": lambda: L.ref_compress(LLVM, P(m), ns, P(idx)),
        "mkperm_4096": lambda: (L.ref_block_mkperm(LLVM, P(keys), nk, nk, BUCKETS, P(perm), P(offs)), L.ref_sync()),
        "scatter_add_f32": scatter,
    }
    times = {k: [] for k in prims}
    for it in range(warmup + steps):
        for name, fn in prims.items():
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times[name].append(dt)
    per = {}
    tot_bytes = tot_time = 0.0
    for name in prims:
        t = sum(times[name]) / len(times[name])
        nbytes = size[name] * bpe[name]
        per[name] = {"elements": size[name], "GBps": nbytes / t / 1e9, "ms": t * 1e3}
        tot_bytes += nbytes; tot_time += t
    per["scatter_add_f32"]["kind"] = ("port: oracle.c restatement of the reference's ReduceMode::Expand scatter "
                                      "(its LLVM JIT is not available); the other six are the unmodified reference")
    return tot_bytes / tot_time / 1e9, per, cores, tot_time


def cpu_sample_text(scale, cores):
    sz = ("the full BASELINE sizes (2^28 f32 / 2^30 u32 / 2^30 u8 / 2^26 keys / 2^28 -> 2^20 bins)" if scale == 0
          else f"the BASELINE sizes shrunk by 2^{scale}")
    return (f"whole suite (7 primitives) at {sz}, host arrays, unmodified reference LLVM-backend primitives "
            f"(oracle/_ref) on {cores} threads; scatter_add through oracle.c's Expand-mode restatement")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, per, cores, step_s = cpu_reference_run(args.steps, max(1, min(args.warmup, 2)), args.scale)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f32/u32/u8",
        "data": "synthetic (fmix32)", "config": workload_config(args, weak=args.weak),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "reference",
                         "sample": cpu_sample_text(args.scale, cores)},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "primitives": per, "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
#  NUMA placement of a rank (pinned staging buffers are first-touched after this)
# ------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Restricts this process to the CPUs of the NUMA node its GPU hangs off, so that pinned
    buffers allocated afterwards are node-local. Returns a description (or why it did nothing)."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return "unknown (nvidia-smi gave no bus id)"
        bus = bus[-12:] if len(bus) > 12 else bus          # 00000000:1B:00.0 -> 0000:1b:00.0
        base = f"/sys/bus/pci/devices/{bus}"
        with open(f"{base}/numa_node") as f:
            node = int(f.read().strip())
        with open(f"{base}/local_cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return f"node {node} (no affinity change)"
        os.sched_setaffinity(0, cpus)
        return f"node {node}, {len(cpus)} cpus"
    except (OSError, ValueError, subprocess.SubprocessError) as e:
        return f"unknown ({type(e).__name__})"


# ------------------------------------------------------------------------------------------
#  GPU arm
# ------------------------------------------------------------------------------------------
class Workload:
    """Device-resident shard of every input plus the output buffers of one rank."""

    def __init__(self, torch, ops, ddist, dev, rank, world, scale, weak):
        self.size = {name: 1 << (lg - scale) for name, lg, _, _ in SUITE}       # global (strong) / per rank (weak)
        self.bins = 1 << max(4, BINS_LOG2 - scale)
        self.weak, self.rank, self.world = weak, rank, world

        def rng(name, align=1):
            n = self.size[name]
            if weak:
                return rank * n, (rank + 1) * n
            b = ddist.shard_bounds(n, world, align)
            return b[rank], b[rank + 1]

        self.range = {"sum_f32": rng("sum_f32", 256), "prefix_sum_u32": rng("prefix_sum_u32", 64),
                      "compress_u8": rng("compress_u8", 64), "mkperm_4096": rng("mkperm_4096", 64),
                      "scatter_add_f32": rng("scatter_add_f32", 64)}
        self.range["block_reduce256_f32"] = self.range["dot_f32"] = self.range["sum_f32"]
        lo, hi = self.range["sum_f32"]
        # shards that approach the L2 size: one input array per primitive (see config.l2)
        self.distinct = (hi - lo) * 4 < 4 * L2_BYTES

        def f32(lo, hi, xor=0):
            t = torch.empty(hi - lo, dtype=torch.float32, device=dev)
            return ops.fill_fmix32(t, 1, start=lo, xor=xor)

        self.x = f32(lo, hi)
        self.x_br = f32(lo, hi) if self.distinct else self.x
        self.x_dot = f32(lo, hi) if self.distinct else self.x
        self.y = f32(lo, hi, xor=0x9E3779B9)
        lo, hi = self.range["prefix_sum_u32"]
        self.u = ops.fill_fmix32(torch.empty(hi - lo, dtype=torch.int32, device=dev), 0, start=lo)
        self.u_out = torch.empty_like(self.u)
        lo, hi = self.range["compress_u8"]
        self.mask = ops.fill_fmix32(torch.empty(hi - lo, dtype=torch.uint8, device=dev), 2, start=lo, and_=128)
        self.c_out = torch.empty(hi - lo, dtype=torch.int32, device=dev)
        lo, hi = self.range["mkperm_4096"]
        self.keys = ops.fill_fmix32(torch.empty(hi - lo, dtype=torch.int32, device=dev), 0, start=lo, and_=BUCKETS - 1)
        self.perm = torch.empty_like(self.keys)
        lo, hi = self.range["scatter_add_f32"]
        self.sidx = ops.fill_fmix32(torch.empty(hi - lo, dtype=torch.int32, device=dev), 0, start=lo,
                                    xor=0x85EBCA6B, and_=self.bins - 1)
        self.sval = f32(lo, hi) if (self.distinct or (lo, hi) != self.range["sum_f32"]) else self.x
        self.bins_t = torch.zeros(self.bins, dtype=torch.float32, device=dev)
        self.br_out = torch.empty((self.x.numel() + 255) // 256, dtype=torch.float32, device=dev)

    def elements(self, name):
        """elements all ranks process together"""
        return self.size[name] * (self.world if self.weak else 1)

    def total_bytes(self):
        return sum(self.elements(name) * b for name, _, b, _ in SUITE)


def make_prims(wl, sh, ops, ReduceOp, VarType, results):
    """The seven primitives of one step as closures (public API of drjit_b200)."""
    world = sh.world
    zero = b"\0\0\0\0"
    import torch
    dev = wl.x.device
    fused = sh.comm is not None
    # preallocated results of the fused path (a per-call torch.empty costs microseconds that show next
    # to 20-100 us primitives at N = 8)
    o_sum = torch.empty(1, dtype=torch.float32, device=dev); o_dot = torch.empty(1, dtype=torch.float32, device=dev)
    o_off = torch.empty(1, dtype=torch.int32, device=dev)
    o_hist = torch.empty(BUCKETS, dtype=torch.int32, device=dev); o_rb = torch.empty(BUCKETS, dtype=torch.int32, device=dev)

    def p_sum():
        results["sum"] = sh.reduce(ReduceOp.Add, wl.x, out=o_sum) if fused else sh.reduce(ReduceOp.Add, wl.x)

    def p_block_reduce():
        results["br"] = sh.block_reduce(ReduceOp.Add, wl.x_br, 256, out=wl.br_out)   # block-aligned shards: no exchange

    def p_dot():
        results["dot"] = sh.dot(wl.x_dot, wl.y, out=o_dot) if fused else sh.dot(wl.x_dot, wl.y)

    def p_prefix():
        if world > 1:   # shard-offset form (see module docstring / DESIGN.md section 5)
            results["scan"] = sh.prefix_reduce_offsets(ReduceOp.Add, wl.u, vt=VarType.UInt32, out=wl.u_out,
                                                       **({"offset": o_off} if fused else {}))
        else:
            results["scan"] = (ops.block_prefix_reduce(ReduceOp.Add, wl.u, wl.u.numel(), True, False,
                                                       vt=VarType.UInt32, out=wl.u_out), None)

    def p_compress():
        if world > 1:
            results["count"] = sh.compress(wl.mask, wl.range["compress_u8"][0] & 0xFFFFFFFF, out=wl.c_out)
        else:           # the seam function (cuda_ts.cpp:683-763): one launch, count read from pinned memory after the wait
            results["count"] = (wl.c_out, [ops.compress_into(wl.mask, wl.c_out)])

    def p_mkperm():
        if world > 1:
            results["mkperm"] = sh.mkperm(wl.keys, BUCKETS, wl.range["mkperm_4096"][0] & 0xFFFFFFFF, perm=wl.perm,
                                          **({"hist": o_hist, "rank_base": o_rb, "raw_table": True} if fused else {}))
        else:           # the seam function jit_var_call_reduce calls (call.cpp:1324): pinned table + count
            results["mkperm"] = ops.block_mkperm(wl.keys, wl.keys.numel(), BUCKETS, perm=wl.perm, raw_table=True)

    def p_scatter():
        ops.memset(wl.bins_t, zero)
        results["bins"] = sh.scatter_add(wl.bins_t, wl.sval, wl.sidx)

    return [("sum_f32", p_sum), ("block_reduce256_f32", p_block_reduce), ("dot_f32", p_dot),
            ("prefix_sum_u32", p_prefix), ("compress_u8", p_compress), ("mkperm_4096", p_mkperm),
            ("scatter_add_f32", p_scatter)]


def verify(torch, dist, wl, results, rank, world, dev):
    """One untimed pass checked against torch-computed invariants on every rank (bit-exact for the
    integer results, 1e-6 * log2 N relative for f32 sums). Returns {primitive: bool}, AND-ed over ranks."""
    ok = {}
    M32 = 0xFFFFFFFF

    def allsum(t):
        if world > 1:
            dist.all_reduce(t)
        return t

    def gather(t):
        if world == 1:
            return t.unsqueeze(0)
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out

    CH = 1 << 24
    # ---- sum / dot (f32, tolerance) and block_reduce
    n_glob = wl.elements("sum_f32")
    tol = 1e-6 * math.log2(max(n_glob, 2))
    exp = float(allsum(torch.sum(wl.x, dtype=torch.float64).reshape(1))[0])
    got = float(results["sum"].cpu()[0])
    ok["sum_f32"] = abs(got - exp) <= tol * abs(exp)
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    for lo in range(0, wl.x_dot.numel(), CH):
        acc += torch.sum(wl.x_dot[lo:lo + CH].double() * wl.y[lo:lo + CH].double())
    exp = float(allsum(acc)[0])
    got = float(results["dot"].cpu()[0])
    ok["dot_f32"] = abs(got - exp) <= tol * abs(exp)
    nb = wl.x_br.numel() // 256
    expb = torch.sum(wl.x_br[:nb * 256].view(nb, 256), dim=1, dtype=torch.float64)
    gotb = results["br"][:nb].double()
    ok["block_reduce256_f32"] = bool(torch.all((gotb - expb).abs() <= 1e-6 * 8 * expb.abs().clamp_min(1.0)))
    del expb, gotb

    # ---- exclusive prefix sum (u32, bit-exact): global[i] = offset + local[i]
    local, offset = results["scan"]
    shard_total = torch.sum(wl.u.long() & M32).reshape(1) & M32
    lower = int(gather(shard_total)[:rank].sum().item()) & M32 if world > 1 else 0
    good = True
    if offset is not None:
        good = (int(offset.cpu()[0]) & M32) == lower
    carry = lower
    for lo in range(0, wl.u.numel(), CH):
        v = wl.u[lo:lo + CH].long() & M32
        inc = torch.cumsum(v, 0)
        expc = (inc - v + carry) & M32
        gotc = ((local[lo:lo + CH].long() & M32) + (lower if offset is not None else 0)) & M32
        good = good and bool(torch.equal(expc, gotc))
        carry = (carry + int(inc[-1].item())) & M32
    ok["prefix_sum_u32"] = good

    # ---- compress: counts of all ranks + exact index list of this shard
    c_out, counts = results["count"]
    base = wl.range["compress_u8"][0]
    my = torch.sum(wl.mask != 0).reshape(1)
    exp_counts = [int(c) for c in gather(my).flatten().cpu().tolist()]
    good = [int(c) for c in counts] == exp_counts
    pos = 0
    for lo in range(0, wl.mask.numel(), CH):
        nz = (torch.nonzero(wl.mask[lo:lo + CH]).flatten() + (lo + base)) & M32
        good = good and bool(torch.equal(c_out[pos:pos + nz.numel()].long() & M32, nz))
        pos += nz.numel()
    ok["compress_u8"] = good and pos == exp_counts[rank]

    # ---- mkperm: permutation of the shard, keys non-decreasing along it, histogram, global table
    res = results["mkperm"]
    kbase = wl.range["mkperm_4096"][0]
    nk = wl.keys.numel()
    hist_exp = torch.bincount(wl.keys.long(), minlength=BUCKETS)
    allh = gather(hist_exp)
    gsize = allh.sum(0)
    gstart = torch.cumsum(gsize, 0) - gsize
    ids = torch.nonzero(gsize).flatten()
    table_exp = torch.stack([ids, gstart[ids], gsize[ids], torch.zeros_like(ids)], 1).cpu()
    if world > 1:
        perm, table = res.perm, res.table
        if isinstance(table, tuple):        # raw pinned table + unique count (fused path)
            table = table[0][:4 * table[1]].view(-1, 4).to(torch.int64) & M32
        good = bool(torch.equal(res.hist.long() & M32, hist_exp))
        good = good and bool(torch.equal(res.rank_base.long() & M32, (gstart + allh[:rank].sum(0)) & M32))
    else:
        perm, offsets, unique = res
        table = (offsets[:4 * unique].view(-1, 4).to(torch.int64) & M32)
        good = True
    p = ((perm.long() & M32) - kbase)
    good = good and bool(p.min() >= 0) and bool(p.max() < nk) and bool(torch.all(torch.bincount(p, minlength=nk) == 1))
    k = wl.keys[p]
    good = good and bool(torch.all(k[1:] >= k[:-1]))
    ok["mkperm_4096"] = good and bool(torch.equal(table.cpu(), table_exp))
    del p, k

    # ---- scatter_add: bins against an f64 accumulation of the global array
    expb = torch.zeros(wl.bins, dtype=torch.float64, device=dev)
    for lo in range(0, wl.sval.numel(), CH):
        expb.index_add_(0, wl.sidx[lo:lo + CH].long(), wl.sval[lo:lo + CH].double())
    allsum(expb)
    ok["scatter_add_f32"] = bool(torch.all((results["bins"].double() - expb).abs() <= 1e-5 * expb.abs().clamp_min(1.0)))

    flags = torch.tensor([int(bool(ok[name])) for name, _, _, _ in SUITE], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    return {name: bool(int(f)) for (name, _, _, _), f in zip(SUITE, flags.cpu().tolist())}


def time_suite(torch, dist, dr, prims, steps, world, dev):
    """Times exactly `steps` steps (barrier + synchronize on both sides, device events, MAX over
    ranks). Returns (ms per step, [ms per primitive], launches, (t0, t1) monotonic)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in prims]
          for _ in range(steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    dr.launch_count(reset=True)
    t0 = time.monotonic()
    start.record()
    for k in range(steps):
        for i, (_, fn) in enumerate(prims):
            ev[k][i][0].record()
            fn()
            ev[k][i][1].record()
    end.record()
    barrier()
    t1 = time.monotonic()
    launches = dr.launch_count()
    total_ms = start.elapsed_time(end)
    per_ms = [sum(ev[k][i][0].elapsed_time(ev[k][i][1]) for k in range(steps)) / steps for i in range(len(prims))]
    t = torch.tensor([total_ms] + per_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / steps, [float(v) for v in t[1:]], launches, (t0, t1)


def primitive_table(wl, per_ms, peak, world):
    out = {}
    for (name, _, bpe, _), ms in zip(SUITE, per_ms):
        el = wl.elements(name)
        gbs = el * bpe / (ms * 1e-3) / 1e9
        out[name] = {"elements": el, "ms": round(ms, 4), "GBps": round(gbs, 1),
                     "Gelem_per_s": round(el / (ms * 1e-3) / 1e9, 2),
                     "frac_of_peak_per_gpu": round(gbs / (peak * world), 4)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200,
                    help="timed steps (default 200: ~1 s, long enough for nvidia-smi to sample the clocks inside the timed region)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=0, help="shrink every array by 2^SCALE (debug)")
    ap.add_argument("--weak", action="store_true", help="weak scaling: every rank owns a full-size shard")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the verified pass (profiler runs only; the JSON says so)")
    ap.add_argument("--no-extras", action="store_true", help="skip the NCCL-path, materialised-scan and weak-scaling side measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "not bound (single rank)"

    import torch
    import torch.distributed as dist

    import drjit_b200 as dr
    from drjit_b200 import ReduceOp, VarType, ops
    from drjit_b200 import dist as ddist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        comm = ddist.PeerComm.from_process_group(dist.group.WORLD, device=dev, bulk_bytes=8 << 20)
    group = dist.group.WORLD if world > 1 else None
    sh = ddist.Sharded(rank=rank, world=world, group=group, comm=comm)          # product path (fused combine)
    sh_nccl = ddist.Sharded(rank=rank, world=world, group=group)                # library-collective baseline

    peak, peak_src = measured_peaks()
    wl = Workload(torch, ops, ddist, dev, rank, world, args.scale, args.weak)
    results = {}
    prims = make_prims(wl, sh, ops, ReduceOp, VarType, results)
    total_bytes = wl.total_bytes()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(pr):
        for _, fn in pr:
            fn()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    barrier()
    t_load = time.monotonic()
    # W warm-up steps, padded (same count on every rank) so that the clocks are sampled under
    # this load for >= 1 s before the timed region starts
    for _ in range(args.warmup):
        one_step(prims)
    barrier()
    elapsed = time.monotonic() - t_load
    extra = 0 if elapsed >= 1.0 else min(2000, int((1.0 - elapsed) / max(elapsed / args.warmup, 1e-4)) + 1)
    if world > 1:
        t = torch.tensor([extra], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        extra = int(t.item())
    for _ in range(extra):
        one_step(prims)

    ms_per_step, per_ms, launches, (t_timed0, t_timed1) = time_suite(torch, dist, dr, prims, args.steps, world, dev)
    clocks = sampler.stop(t_load, t_timed0, t_timed1) if rank == 0 else None
    value = total_bytes / (ms_per_step * 1e-3) / 1e9
    primitives = primitive_table(wl, per_ms, peak, world)

    # ---- one untimed, verified pass on every rank ------------------------------------------
    checked, verified = {}, None
    if not args.no_verify:
        one_step(prims)
        barrier()
        checked = verify(torch, dist, wl, results, rank, world, dev)
        verified = all(checked.values())
        if not verified:
            raise SystemExit(f"bench.py: verification failed on the {world}-rank run: {checked}")

    # ---- side measurements (NOT part of `value`) --------------------------------------------
    side_steps = max(3, min(args.steps, 20))
    weak_scaling = None
    if world > 1 and not args.no_extras:
        # (1) the same suite with the combine steps over NCCL collectives instead of peer memory
        res2 = {}
        pr2 = make_prims(wl, sh_nccl, ops, ReduceOp, VarType, res2)
        one_step(pr2)
        ms2, per2, _, _ = time_suite(torch, dist, dr, pr2, side_steps, world, dev)
        for (name, _, _, _), ms in zip(SUITE, per2):
            primitives[name]["nccl_path_ms"] = round(ms, 4)
        primitives["_suite"] = {"fused_ms": round(ms_per_step, 4), "nccl_path_ms": round(ms2, 4)}
        # (2) the scan in materialised form (every element carries the global value: 12 B/element)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sh.prefix_sum(wl.u, vt=VarType.UInt32, out=wl.u_out)
        barrier()
        a.record()
        for _ in range(side_steps):
            sh.prefix_sum(wl.u, vt=VarType.UInt32, out=wl.u_out)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / side_steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        gbs = wl.elements("prefix_sum_u32") * 8.0 / (ms * 1e-3) / 1e9
        primitives["prefix_sum_u32"]["form"] = "shard-offset (local scan + per-shard offset)"
        primitives["prefix_sum_u32"]["materialised_form"] = {
            "ms": round(ms, 4), "GBps": round(gbs, 1), "frac_of_peak_per_gpu": round(gbs / (peak * world), 4),
            "note": "global value in every element (extra read pass over the shard); informational, not in `value`"}

    # ---- end-to-end: host buffers through the public API, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, ops, sh, dev, world, rank, wl, prims, results, total_bytes)
        if e2e is not None:
            e2e["numa"] = numa

    # ---- weak scaling beside the strong-scaling headline -------------------------------------
    if world > 1 and not args.weak and not args.no_extras:
        del wl, prims
        results.clear()
        torch.cuda.empty_cache()
        wlw = Workload(torch, ops, ddist, dev, rank, world, args.scale, True)
        resw = {}
        prw = make_prims(wlw, sh, ops, ReduceOp, VarType, resw)
        for _ in range(3):
            one_step(prw)
        msw, perw, _, _ = time_suite(torch, dist, dr, prw, side_steps, world, dev)
        weak_scaling = {"value": round(wlw.total_bytes() / (msw * 1e-3) / 1e9, 1), "unit": "GB/s", "ms_per_step": round(msw, 4),
                        "steps": side_steps, "note": f"every rank owns a full-size shard (global arrays {world}x larger)",
                        "primitives_ms": {name: round(ms, 4) for (name, _, _, _), ms in zip(SUITE, perw)}}
        wl_distinct = False
    else:
        wl_distinct = wl.distinct

    if rank != 0:
        if world > 1:
            dist.barrier()
            comm.destroy()
            dist.destroy_process_group()
        return

    # per-primitive rooflines; `roofline` = the kernel with the largest share of the step
    rooflines = []
    for (name, lg, bpe, kernel), ms in zip(SUITE, per_ms):
        p = primitives[name]
        per_gpu_elems = p["elements"] // world
        rooflines.append({"primitive": name, "kernel": kernel, "bound": "hbm",
                          "achieved": round(p["GBps"] / world, 1), "peak": peak, "unit": "GB/s",
                          "frac": round(p["GBps"] / world / peak, 4),
                          "traffic": measured_traffic(name, per_gpu_elems),
                          "algorithmic_bytes_per_launch": per_gpu_elems * bpe,
                          "share_of_step": round(ms / sum(per_ms), 4)})
    roofline = dict(max(rooflines, key=lambda r: r["share_of_step"]))
    roofline["peak_source"] = peak_src
    roofline["note"] = ("dominant kernel = largest share of the step's device time (CUDA events around every primitive "
                        "call); all seven are listed in `rooflines`")

    cpu_baseline = None
    if not args.no_cpu:
        try:
            v, per, cores, _ = cpu_reference_run(2, 1, args.scale)
            cpu_baseline = {"value": round(v, 2), "unit": "GB/s", "cores": cores, "kind": "reference",
                            "sample": cpu_sample_text(args.scale, cores) + "; 2 timed steps after 1 warm-up",
                            "primitives": {k: round(p["GBps"], 2) for k, p in per.items()}}
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference",
                            "sample": f"unavailable: {e}"}

    scatter_forms = None
    if world == 1 and not args.no_extras:
        try:
            del wl, prims
            results.clear()
            torch.cuda.empty_cache()
            scatter_forms = time_scatter_forms(torch, ops, dev, peak, args.scale)
        except Exception as e:  # pragma: no cover  (a side measurement never takes the bench line with it)
            scatter_forms = {"unavailable": str(e)}

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f32/u32/u8",
        "data": "synthetic (fmix32)", "config": workload_config(args, weak=args.weak, distinct=wl_distinct),
        "frac_of_hbm_peak": round(value / (peak * world), 4), "verified": verified, "verified_primitives": checked,
        "primitives": primitives, "roofline": roofline, "rooflines": rooflines, "cpu_baseline": cpu_baseline,
        "e2e": e2e, "weak_scaling": weak_scaling, "scatter_forms": scatter_forms, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        comm.destroy()
        dist.destroy_process_group()


def time_scatter_forms(torch, ops, dev, peak, scale):
    """Side measurement (N = 1, NOT part of `value`): the packet form of scatter-reduce and dr.scatter_inc
    (drjit_b200/csrc/scatter_packet.cu), CUDA events around 10 calls each after 3 warm-ups."""
    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    out = {}
    n, pixels = (1 << 26) >> scale, max((1 << 20) >> scale, 16)
    vals = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(4)]
    for k, v in enumerate(vals):
        ops.fill_fmix32(v, 1, xor=k + 1)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    ops.fill_fmix32(idx, 0, xor=0x85EBCA6B, and_=pixels - 1)
    film = torch.zeros(4 * pixels, dtype=torch.float32, device=dev)
    ms = timed(lambda: ops.scatter_reduce_packet(1, film, vals, idx))
    out["scatter_add_packet4_f32"] = {"elements": n, "ms": round(ms, 4), "GBps": round(n * 20 / ms / 1e6, 1),
                                      "frac_of_peak_per_gpu": round(n * 20 / ms / 1e6 / peak, 4),
                                      "note": "2^26 four-component packets into 2^20 packets (20 B/element); bound by the L2's 16-byte reduction rate"}
    del vals, film
    n = (1 << 28) >> scale
    mask = torch.empty(n, dtype=torch.uint8, device=dev)
    ops.fill_fmix32(mask, 2, and_=128)
    slots = torch.empty(n, dtype=torch.int32, device=dev)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    ms = timed(lambda: ops.scatter_inc(counter, None, active=mask, size=n, out=slots))
    out["scatter_inc_queue_masked"] = {"elements": n, "ms": round(ms, 4), "GBps": round(n * 5 / ms / 1e6, 1),
                                       "frac_of_peak_per_gpu": round(n * 5 / ms / 1e6 / peak, 4),
                                       "note": "dr.scatter_inc(counter, 0, active) over 2^28 elements, 50 % active (mask byte + slot: 5 B/element)"}
    return out


def run_e2e(args, torch, dist, ops, sh, dev, world, rank, wl, prims, results, total_bytes):
    """Same suite, but every step first copies the step's inputs host->device from pinned
    memory and afterwards reads every primitive's result back to the host."""
    from drjit_b200 import ReduceOp, VarType
    # a user uploads x once and runs sum / block_reduce / dot on it (uploads come between the
    # primitives here, so L2 residency is not a concern as it is in the device-timed run)
    saved = (wl.x_br, wl.x_dot)
    wl.x_br = wl.x_dot = wl.x
    inputs = dict(x=wl.x, y=wl.y, u=wl.u, mask=wl.mask, keys=wl.keys, sidx=wl.sidx, sval=wl.sval)
    outputs = dict(u_out=wl.u_out, c_out=wl.c_out, perm=wl.perm, bins_t=wl.bins_t, br_out=wl.br_out)
    host_in = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in inputs.items() if k != "sval" or v is not wl.x}
    for k, h in host_in.items():
        h.copy_(inputs[k])
    host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in outputs.items()}
    h2d = sum(h.numel() * h.element_size() for h in host_in.values())
    steps = max(2, args.steps)

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
    offset = torch.zeros(1, dtype=u_dev.dtype, device=dev)

    # Downloads of step k overlap the uploads of step k + 1 (both directions of the link stay busy
    # across the step boundary): a kernel that overwrites a result buffer first waits for the
    # previous step's download of that buffer; the host waits for the downloads once, at the end of
    # the timed region. Scalars are still read synchronously in every step.
    dl_done = {}

    def scan_chunked(up_u):
        moved = 0
        for c, (lo, hi) in enumerate(bounds):
            main.wait_event(up_u[c])
            if ("u_out", c) in dl_done:
                main.wait_event(dl_done[("u_out", c)])      # the previous step's download of this piece
            ops.prefix_reduce_carry(ReduceOp.Add, u_dev[lo:hi], True, False, carry_in=carry[c & 1] if c else None,
                                    total_out=carry[(c + 1) & 1], vt=VarType.UInt32, out=u_res[lo:hi])
            done = torch.cuda.Event(); done.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                host_out["u_out"][lo:hi].copy_(u_res[lo:hi], non_blocking=True)
                e = torch.cuda.Event(); e.record(s_out); dl_done[("u_out", c)] = e
            moved += (hi - lo) * u_res.element_size()
        if world > 1:       # shard-offset form: exclusive fold of the shard totals over the lower ranks
            sh.fold_scalar(ReduceOp.Add, carry[len(bounds) & 1], offset, lower=True, vt=VarType.UInt32)
            results["scan"] = (u_res, offset)
        else:
            results["scan"] = (u_res, None)
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
            for k in outs:
                if k in dl_done:
                    main.wait_event(dl_done[k])
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
                    e = torch.cuda.Event(); e.record(s_out); dl_done[k] = e
        for k in ("sum", "dot"):
            results[k].cpu(); d2h += 4
        if results["scan"][1] is not None:          # shard-offset form: the offset travels with the scan
            results["scan"][1].cpu(); d2h += 4
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
    t = torch.tensor([dt, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt = float(tmax[0])
    h2d_all, d2h_all = int(t[1]), int(t[2])
    wl.x_br, wl.x_dot = saved
    return {"value": round(total_bytes / dt / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d_all,
            "d2h_bytes_per_step": d2h_all, "ms_per_step": round(dt * 1e3, 3), "steps": steps,
            "pcie_GBps": {"h2d": round(h2d_all / dt / 1e9, 1), "d2h": round(d2h_all / dt / 1e9, 1),
                          "note": "all ranks together, per direction, averaged over the step"},
            "note": "per rank: pinned host inputs -> device, suite through the public API, every result "
                    "(scalars, block sums, scan, index list, permutation, bins) -> pinned host; uploads, "
                    "kernels and downloads pipelined on three streams (downloads of a step overlap the uploads of "
                    "the next; all copies complete inside the timed region), the scan fed in 8 pieces through its carry form"}


if __name__ == "__main__":
    main()
