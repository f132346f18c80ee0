#!/usr/bin/env python
"""Incumbent beside scatter_packet.cu: the UNMODIFIED reference's CUDA JIT (oracle/_ref, through
oracle/ref_build/ref_shim.cpp) running dr.scatter_add of a 4-component packet and dr.scatter_inc on the
same B200, timed with the host clock around call + jit_sync_thread (kernels of 0.3 ms and more; the
JIT-compiled kernel is cached by the warm-up call). Developer tool; not the bench contract.

    python scripts/time_scatter_ref.py [--log2 26]
"""
import argparse
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drjit_b200 as dr  # noqa: E402
from drjit_b200 import ops  # noqa: E402
from oracle import ref  # noqa: E402
from oracle.capi import OP, VT  # noqa: E402

vp = ctypes.c_void_p


def best_ms(fn, sync, reps=5):
    fn(); sync()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); sync(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


KERNEL_HISTORY = 1 << 15        # JitFlag::KernelHistory, jit.h:1734


def ref_kernel_ms(L, fn):
    """Device time of the JIT-compiled kernel(s) of one reference call: the largest execution_time of
    jit_kernel_history() (jit.h:2700-2736) -- the host clock around the call also contains the tracer."""
    N = 64
    b = (ctypes.c_uint32 * N)(); t = (ctypes.c_uint32 * N)(); s = (ctypes.c_uint32 * N)(); ms = (ctypes.c_float * N)()
    fn(); L.ref_sync()
    L.ref_set_flag(KERNEL_HISTORY, 1); L.ref_kernel_history_clear()
    best = None
    for _ in range(3):
        fn(); L.ref_sync()
        cnt = L.ref_kernel_history(b, t, s, ms, N)
        jit = [ms[i] for i in range(min(cnt, N)) if t[i] == 0]          # KernelType::JIT
        if jit:
            best = max(jit) if best is None else min(best, max(jit))
    L.ref_set_flag(KERNEL_HISTORY, 0)
    return best


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--log2", type=int, default=26); a = ap.parse_args()
    n = 1 << a.log2
    L = ref.lib(cuda=True, llvm=False)
    if not ref.has_backend(ref.CUDA) or not hasattr(L, "ref_scatter_packet"):
        print("reference CUDA backend not available"); return
    dev = "cuda"
    sync = torch.cuda.synchronize
    # ---- packet scatter-add: 2^log2 RGBA samples into 2^20 pixels
    count, pixels = 4, 1 << 20
    vals = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(count)]
    for k, v in enumerate(vals):
        ops.fill_fmix32(v, 1, xor=k + 1)
    idx = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(idx, 0, xor=0x85EBCA6B, and_=pixels - 1)
    tgt = torch.zeros(count * pixels, dtype=torch.float32, device=dev)
    ptrs = (vp * count)(*[v.data_ptr() for v in vals])
    sync()
    for mode, name in ((1, "Direct"), (2, "Local")):
        t = best_ms(lambda: L.ref_scatter_packet(ref.CUDA, VT["f32"], OP["add"], mode, vp(tgt.data_ptr()), pixels, ptrs, count,
                                                 vp(idx.data_ptr()), None, n), L.ref_sync)
        k = ref_kernel_ms(L, lambda: L.ref_scatter_packet(ref.CUDA, VT["f32"], OP["add"], mode, vp(tgt.data_ptr()), pixels, ptrs, count,
                                                          vp(idx.data_ptr()), None, n))
        print(f"reference JIT  packet4 {name:6s} n=2^{a.log2}  {t:8.3f} ms  {n / t / 1e6:8.2f} Gelem/s  kernel (KernelHistory) {k} ms", flush=True)
    t = best_ms(lambda: dr.scatter_add(tgt, vals, idx), sync)
    print(f"drjit_b200     packet4        n=2^{a.log2}  {t:8.3f} ms  {n / t / 1e6:8.2f} Gelem/s  (host clock, same method)", flush=True)
    # ---- scatter_inc: queue counter (all indices 0), 16 random counters
    out = torch.empty(n, dtype=torch.int32, device=dev)
    for B, name in ((1, "queue"), (16, "16 counters")):
        ctr = torch.zeros(B, dtype=torch.int32, device=dev)
        ops.fill_fmix32(idx, 0, xor=7, and_=B - 1)
        sync()
        t = best_ms(lambda: L.ref_scatter_inc(ref.CUDA, vp(ctr.data_ptr()), B, vp(idx.data_ptr()), None, n, vp(out.data_ptr())), L.ref_sync)
        k = ref_kernel_ms(L, lambda: L.ref_scatter_inc(ref.CUDA, vp(ctr.data_ptr()), B, vp(idx.data_ptr()), None, n, vp(out.data_ptr())))
        print(f"reference JIT  scatter_inc {name:12s} n=2^{a.log2}  {t:8.3f} ms host clock (tracer + a 4n-byte copy of the result included)  "
              f"kernel (KernelHistory) {k} ms = {n / k / 1e6 if k else 0:8.2f} Gelem/s", flush=True)
        t = best_ms(lambda: ops.scatter_inc(ctr, None if B == 1 else idx, size=n, out=out), sync)
        print(f"drjit_b200     scatter_inc {name:12s} n=2^{a.log2}  {t:8.3f} ms  {n / t / 1e6:8.2f} Gelem/s", flush=True)


if __name__ == "__main__":
    main()
