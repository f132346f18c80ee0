#!/usr/bin/env python
"""Per-call cost at wavefront-sized arrays (2^10 .. 2^24 elements): this library next to the
reference's own CUDA kernels on the same GPU (what limits small-array
throughput in Mitsuba's wavefront loops is launch count and scratch handling, not bandwidth).

Both sides are driven through their C entry points with ctypes and preallocated device buffers;
each figure is the wall time of 200 back-to-back calls plus one final synchronisation, divided by
200 (asynchronous primitives), or of 50 calls for the synchronous ones (compress, mkperm + table).

    python scripts/small_sizes.py
"""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drjit_b200 as dr  # noqa: E402
from drjit_b200 import ops  # noqa: E402
from drjit_b200._lib import check, lib  # noqa: E402
from oracle import ref  # noqa: E402
from oracle.capi import OP, VT  # noqa: E402

vp = ctypes.c_void_p


def per_call(fn, sync, calls):
    for _ in range(3):
        fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(calls):
        fn()
    sync()
    return (time.perf_counter() - t0) / calls * 1e6


def main():
    L = ref.lib(cuda=True, llvm=False)
    have_ref = ref.has_backend(ref.CUDA)
    stream = vp(torch.cuda.current_stream().cuda_stream)
    print(f"{'primitive':22s} {'n':>6s} {'reference CUDA':>16s} {'this library':>14s}   (us per call)")
    for lg in range(10, 25, 2):
        n = 1 << lg
        x = torch.empty(n, dtype=torch.int32, device="cuda"); ops.fill_fmix32(x, 0)
        out = torch.empty_like(x)
        one = torch.empty(1, dtype=torch.int32, device="cuda")
        mbuf = torch.zeros(n + 4096, dtype=torch.uint8, device="cuda"); m = mbuf[:n]; ops.fill_fmix32(m, 2, and_=128)
        keys = torch.empty(n, dtype=torch.int32, device="cuda"); ops.fill_fmix32(keys, 0, and_=63)
        torch.cuda.synchronize()
        B = 64
        off_o = torch.empty(4 * B + 1, dtype=torch.int32).pin_memory()
        off_r = L.ref_malloc(ref.CUDA, 4 * (4 * B + 1), 1) if have_ref else None
        cnt = ctypes.c_uint32(0)

        cases = [
            ("sum u32", 200,
             lambda: L.ref_block_reduce(ref.CUDA, VT["u32"], OP["add"], n, n, vp(x.data_ptr()), vp(one.data_ptr())),
             lambda: check(lib.drjit_b200_block_reduce(stream, VT["u32"], OP["add"], n, n, vp(x.data_ptr()), vp(one.data_ptr())))),
            ("exclusive prefix u32", 200,
             lambda: L.ref_block_prefix_reduce(ref.CUDA, VT["u32"], OP["add"], n, n, 1, 0, vp(x.data_ptr()), vp(out.data_ptr())),
             lambda: check(lib.drjit_b200_block_prefix_reduce(stream, VT["u32"], OP["add"], n, n, 1, 0, vp(x.data_ptr()), vp(out.data_ptr())))),
            ("compress (sync)", 50,
             lambda: L.ref_compress(ref.CUDA, vp(m.data_ptr()), n, vp(out.data_ptr())),
             lambda: check(lib.drjit_b200_compress(stream, vp(m.data_ptr()), n, vp(out.data_ptr()), ctypes.byref(cnt)))),
            ("mkperm 64 + table", 50,
             lambda: L.ref_block_mkperm(ref.CUDA, vp(keys.data_ptr()), n, n, B, vp(out.data_ptr()), vp(off_r)),
             lambda: check(lib.drjit_b200_block_mkperm(stream, vp(keys.data_ptr()), n, n, B, vp(out.data_ptr()),
                                                       vp(off_o.data_ptr()), ctypes.byref(cnt)))),
        ]
        for name, calls, f_ref, f_ours in cases:
            t_ref = per_call(f_ref, L.ref_sync, calls) if have_ref else float("nan")
            t_ours = per_call(f_ours, torch.cuda.synchronize, calls)
            print(f"{name:22s} 2^{lg:<4d} {t_ref:16.1f} {t_ours:14.1f}", flush=True)
        if have_ref:
            L.ref_free(off_r)


if __name__ == "__main__":
    main()
