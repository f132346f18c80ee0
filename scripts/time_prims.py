#!/usr/bin/env python
"""Times individual primitives with CUDA events (developer tool; not the bench contract).

    python scripts/time_prims.py [prim ...] [--log2 N] [--reps R]
prims: sum block_reduce dot scan scan64 scanseg compress compress01 compress99 mkperm mkperm256 scatter sort sortkeys
       packet scatter_inc sort_composed torch_sort (the last two only when named) all
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drjit_b200 as dr  # noqa: E402
from drjit_b200 import ReduceOp, VarType, ops  # noqa: E402

def _peak():
    import json
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        return 6650.0   # B200_PROFILING.md fallback


PEAK = _peak()


WARM = 3


def timeit(fn, reps):
    for _ in range(WARM):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("prims", nargs="*", default=["all"])
    ap.add_argument("--log2", type=int, default=None)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3)
    a = ap.parse_args()
    global WARM
    WARM = a.warm
    want = set(a.prims)
    dev = "cuda"

    def run(name, log2, bytes_per_elem, setup, group=None):
        if "all" not in want and name not in want and group not in want:
            return
        n = 1 << (a.log2 or log2)
        fn = setup(n)
        med, best = timeit(fn, a.reps)
        gbs = n * bytes_per_elem / med / 1e6
        print(f"{name:12s} n=2^{(a.log2 or log2):2d}  median {med:8.3f} ms  best {best:8.3f} ms  "
              f"{gbs:8.1f} GB/s  {gbs / PEAK * 100:5.1f}% of measured peak  {n / med / 1e6:8.2f} Gelem/s", flush=True)

    def s_sum(n):
        x = torch.empty(n, dtype=torch.float32, device=dev); ops.fill_fmix32(x, 1)
        return lambda: dr.sum(x)

    def s_br(n):
        x = torch.empty(n, dtype=torch.float32, device=dev); ops.fill_fmix32(x, 1)
        out = torch.empty(n // 256, dtype=torch.float32, device=dev)
        return lambda: ops.block_reduce(ReduceOp.Add, x, 256, out=out)

    def s_dot(n):
        x = torch.empty(n, dtype=torch.float32, device=dev); ops.fill_fmix32(x, 1)
        y = torch.empty(n, dtype=torch.float32, device=dev); ops.fill_fmix32(y, 1, xor=5)
        return lambda: dr.dot(x, y)

    def s_scan(n):
        x = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(x, 0)
        out = torch.empty_like(x)
        return lambda: ops.block_prefix_reduce(ReduceOp.Add, x, n, True, False, vt=VarType.UInt32, out=out)

    def s_scan64(n):
        x = torch.empty(n, dtype=torch.int64, device=dev); x.view(torch.int32)[::2] = 1
        out = torch.empty_like(x)
        return lambda: ops.block_prefix_reduce(ReduceOp.Add, x, n, True, False, vt=VarType.UInt64, out=out)

    def s_scanseg(n):
        x = torch.empty(n, dtype=torch.float32, device=dev); ops.fill_fmix32(x, 1)
        out = torch.empty_like(x)
        return lambda: ops.block_prefix_reduce(ReduceOp.Add, x, 1000, True, False, out=out)

    def s_compress(thr):
        def setup(n):
            m = torch.empty(n, dtype=torch.uint8, device=dev); ops.fill_fmix32(m, 2, and_=thr)
            out = torch.empty(n, dtype=torch.int32, device=dev)
            cnt = torch.zeros(1, dtype=torch.int32, device=dev)
            return lambda: ops.compress_async(m, 0, out=out, count=cnt)
        return setup

    def s_mkperm(buckets):
        def setup(n):
            k = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(k, 0, and_=buckets - 1)
            perm = torch.empty_like(k); hist = torch.empty(buckets, dtype=torch.int32, device=dev)
            return lambda: ops.mkperm_sharded(k, buckets, 0, perm=perm, hist=hist)
        return setup

    def s_scatter(n):
        v = torch.empty(n, dtype=torch.float32, device=dev); ops.fill_fmix32(v, 1)
        i = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(i, 0, xor=0x85EBCA6B, and_=(1 << 20) - 1)
        bins = torch.zeros(1 << 20, dtype=torch.float32, device=dev)
        return lambda: dr.scatter_add(bins, v, i)

    def s_sort(payload):
        def setup(n):
            k = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(k, 0)
            return (lambda: ops.sort_with_indices(k, vt=VarType.UInt32)) if payload else (lambda: ops.sort(k, vt=VarType.UInt32))
        return setup

    def s_sort_composed(n):
        """the reference's _radix_sort structure (drjit/__init__.py:1698-1772) on this library's stable
        block_mkperm: per pass a digit kernel, block_mkperm(256 buckets), one gather per carried array"""
        k = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(k, 0)

        def fn():
            o, idx = k, torch.arange(n, dtype=torch.int32, device=dev)
            for shift in (0, 8, 16, 24):
                digit = (o >> shift) & 255
                perm, _ = ops.block_mkperm(digit, n, 256, want_offsets=False)
                pl = perm.long()
                o, idx = o[pl], idx[pl]
            return o, idx
        return fn

    def s_torch_sort(n):
        k = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(k, 0)
        return lambda: torch.sort(k, stable=True)

    run("sum", 28, 4, s_sum)
    run("block_reduce", 28, 4 + 4 / 256, s_br)
    run("dot", 28, 8, s_dot)
    run("scan", 30, 8, s_scan)
    run("scan64", 28, 16, s_scan64)
    run("scanseg", 28, 8, s_scanseg)
    run("compress", 30, 3, s_compress(128))
    run("compress01", 30, 1 + 4 * 3 / 256, s_compress(3))
    run("compress10", 30, 1 + 4 * 26 / 256, s_compress(26))
    run("compress99", 30, 1 + 4 * 253 / 256, s_compress(253))
    run("mkperm", 26, 12, s_mkperm(4096))
    run("mkperm256", 26, 12, s_mkperm(256))
    run("scatter", 28, 8, s_scatter)
    def s_call_reduce(npay):
        def setup(n):
            k = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(k, 0, and_=4095)
            pays = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(npay)]
            for i, p in enumerate(pays):
                ops.fill_fmix32(p, 1, xor=i + 1)
            return lambda: ops.call_reduce(k, 4096, pays)
        return setup

    def s_call_reduce_unfused(n):
        k = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(k, 0, and_=4095)
        pays = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]

        def fn():
            perm, table = ops.block_mkperm(k, n, 4096)
            pl = perm.long()
            return [p[pl] for p in pays]
        return fn

    if "call_reduce" in want or "all" in want:
        run("call_reduce", 26, 12, s_call_reduce(0))
        run("call_reduce", 26, 12 + 2 * 8, s_call_reduce(2))
        run("call_reduce", 26, 12 + 2 * 8, s_call_reduce_unfused)
    run("sort", 26, 4 * 20 + 0, s_sort(True))          # key + index: 20 B per element and pass
    run("sortkeys", 26, 4 * 12, s_sort(False))
    # ---- scatter_packet.cu: packet scatter-add (film accumulation) and scatter_inc ------------------------
    def s_packet(count, log2_packets, four_scalar=False):
        def setup(n):
            vals = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(count)]
            for k, v in enumerate(vals):
                ops.fill_fmix32(v, 1, xor=k + 1)
            i = torch.empty(n, dtype=torch.int32, device=dev)
            ops.fill_fmix32(i, 0, xor=0x85EBCA6B, and_=(1 << log2_packets) - 1)
            tgt = torch.zeros(count << log2_packets, dtype=torch.float32, device=dev)
            if not four_scalar:
                return lambda: dr.scatter_add(tgt, vals, i)
            # what a caller without the packet form does: one scalar scatter per component at index * count + k
            idxs = [(i * count + k) for k in range(count)]
            return lambda: [dr.scatter_add(tgt, vals[k], idxs[k]) for k in range(count)]
        return setup

    def s_inc(log2_counters, queue=False, masked=False):
        def setup(n):
            B = 1 << log2_counters
            tgt = torch.zeros(B, dtype=torch.int32, device=dev)
            out = torch.empty(n, dtype=torch.int32, device=dev)
            if queue:
                m = None
                if masked:
                    m = torch.empty(n, dtype=torch.uint8, device=dev); ops.fill_fmix32(m, 2, and_=128)
                return lambda: ops.scatter_inc(tgt, None, active=m, size=n, out=out)
            i = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(i, 0, xor=7, and_=B - 1)
            return lambda: ops.scatter_inc(tgt, i, out=out)
        return setup

    run("packet4", 26, 20, s_packet(4, 20), group="packet")                      # 2^26 RGBA samples -> 2^20 pixels
    run("packet4x1", 26, 32, s_packet(4, 20, four_scalar=True), group="packet")  # the same as four scalar scatters
    run("packet2", 26, 12, s_packet(2, 20), group="packet")
    run("packet8", 26, 36, s_packet(8, 20), group="packet")
    run("inc_queue", 28, 4, s_inc(0, queue=True), group="scatter_inc")                # one counter, no index array
    run("inc_queue_m", 28, 5, s_inc(0, queue=True, masked=True), group="scatter_inc") # the same with a 50 % mask
    run("inc_16", 28, 8, s_inc(4), group="scatter_inc")
    run("inc_2048", 28, 8, s_inc(11), group="scatter_inc")
    run("inc_2^20", 26, 8, s_inc(20), group="scatter_inc")
    if "sort_composed" in want:
        run("sort_composed", 26, 4 * 20, s_sort_composed)
    if "torch_sort" in want:
        run("torch_sort", 26, 4 * 20, s_torch_sort)


if __name__ == "__main__":
    main()
