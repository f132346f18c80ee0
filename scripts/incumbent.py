#!/usr/bin/env python
"""The reference's own CUDA kernels ("the incumbent") next to this library on the same GPU.

Developer / evidence tool ("incumbent beside it"). The UNMODIFIED reference
drjit-core built into oracle/_ref exposes its CUDA backend through the driver API only, so it runs
on the GPU box as is (its PTX is JIT-compiled by the driver). For every BASELINE configuration this
script
  * times the reference primitive (wall clock around call + jit_sync_thread, median of 5) and ours
    (CUDA events, median of 5) on the same device buffers, and
  * checks the integer outputs of both against each other bit for bit (scan, compress index list,
    mkperm bucket table and per-bucket grouping), f32 results within 1e-6 * log2(N).

    python scripts/incumbent.py [--log2-shift S]      # S shrinks every size by 2^S
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drjit_b200 as dr  # noqa: E402
from drjit_b200 import ReduceOp, VarType, ops  # noqa: E402
from oracle import ref  # noqa: E402
from oracle.capi import OP, VT  # noqa: E402


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def time_ref(L, fn, reps=5):
    fn(); L.ref_sync()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); L.ref_sync(); ts.append((time.perf_counter() - t0) * 1e3)
    return median(ts)


def time_ours(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-shift", type=int, default=0)
    ap.add_argument("--buckets", type=int, default=4096, help="bucket count of the block_mkperm row")
    a = ap.parse_args()
    S = a.log2_shift
    L = ref.lib(cuda=True, llvm=False)
    if not ref.has_backend(ref.CUDA):
        print("reference CUDA backend did not initialise on this box"); return
    dev = "cuda"
    vp = ctypes.c_void_p
    rows = []

    def report(name, n, bpe, t_ref, t_ours, check):
        rows.append((name, n, t_ref, t_ours, check))
        print(f"{name:28s} n=2^{int(np.log2(n)):2d}  reference CUDA {t_ref:8.3f} ms ({n * bpe / t_ref / 1e6:7.1f} GB/s)   "
              f"this library {t_ours:8.3f} ms ({n * bpe / t_ours / 1e6:7.1f} GB/s)   x{t_ref / t_ours:5.2f}   {check}", flush=True)

    # ---- sum / block_reduce(256) / dot, f32 ---------------------------------------------------
    n = 1 << (28 - S)
    x = torch.empty(n, dtype=torch.float32, device=dev); ops.fill_fmix32(x, 1)
    y = torch.empty(n, dtype=torch.float32, device=dev); ops.fill_fmix32(y, 1, xor=0x9E3779B9)
    out_r = torch.zeros(1, dtype=torch.float32, device=dev)
    t_ref = time_ref(L, lambda: L.ref_block_reduce(ref.CUDA, VT["f32"], OP["add"], n, n, vp(x.data_ptr()), vp(out_r.data_ptr())))
    res = {}
    t_ours = time_ours(lambda: res.__setitem__("v", dr.sum(x)))
    rel = abs(float(res["v"]) - float(out_r)) / abs(float(out_r))
    report("sum f32", n, 4, t_ref, t_ours, f"rel diff {rel:.1e} (tol {1e-6 * np.log2(n):.1e})")

    br_r = torch.zeros(n // 256, dtype=torch.float32, device=dev)
    t_ref = time_ref(L, lambda: L.ref_block_reduce(ref.CUDA, VT["f32"], OP["add"], n, 256, vp(x.data_ptr()), vp(br_r.data_ptr())))
    br_o = torch.empty_like(br_r)
    t_ours = time_ours(lambda: ops.block_reduce(ReduceOp.Add, x, 256, out=br_o))
    rel = float(((br_o - br_r).abs() / br_r.abs().clamp(min=1)).max())
    report("block_reduce(Add,256) f32", n, 4 + 4 / 256, t_ref, t_ours, f"max rel diff {rel:.1e}")

    host = np.zeros(1, np.float32)
    t_ref = time_ref(L, lambda: L.ref_reduce_dot(ref.CUDA, VT["f32"], vp(x.data_ptr()), vp(y.data_ptr()), n, host.ctypes.data_as(vp)))
    t_ours = time_ours(lambda: res.__setitem__("d", dr.dot(x, y)))
    rel = abs(float(res["d"]) - float(host[0])) / abs(float(host[0]))
    report("dot f32 (ref: var-level API)", n, 8, t_ref, t_ours, f"rel diff {rel:.1e}")
    del x, y, br_r, br_o

    # ---- exclusive prefix sum, u32 ----------------------------------------------------------------
    n = 1 << (30 - S)
    u = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(u, 0)
    o_r = torch.empty_like(u); o_o = torch.empty_like(u)
    t_ref = time_ref(L, lambda: L.ref_block_prefix_reduce(ref.CUDA, VT["u32"], OP["add"], n, n, 1, 0, vp(u.data_ptr()), vp(o_r.data_ptr())))
    t_ours = time_ours(lambda: ops.block_prefix_reduce(ReduceOp.Add, u, n, True, False, vt=VarType.UInt32, out=o_o))
    report("exclusive prefix_sum u32", n, 8, t_ref, t_ours, "bit-exact" if torch.equal(o_r, o_o) else "MISMATCH")
    del u, o_r, o_o

    # ---- compress, 50 % -----------------------------------------------------------------------------
    n = 1 << (30 - S)
    mbuf = torch.zeros(n + 4096, dtype=torch.uint8, device=dev)      # (the reference zero-pads up to a multiple of 2048)
    m = mbuf[:n]; ops.fill_fmix32(m, 2, and_=128)
    c_r = torch.empty(n, dtype=torch.int32, device=dev)
    cnt = {}
    t_ref = time_ref(L, lambda: cnt.__setitem__("r", L.ref_compress(ref.CUDA, vp(m.data_ptr()), n, vp(c_r.data_ptr()))))
    t_ours = time_ours(lambda: cnt.__setitem__("o", dr.compress(m)))
    same = cnt["r"] == cnt["o"].numel() and torch.equal(c_r[:cnt["r"]], cnt["o"].view(torch.int32))
    report("compress (50 %)", n, 3, t_ref, t_ours, f"count {cnt['r']}, index list " + ("bit-exact" if same else "MISMATCH"))
    del mbuf, m, c_r, cnt

    # ---- block_mkperm, 4096 buckets --------------------------------------------------------------------
    n = 1 << (26 - S); B = a.buckets
    keys = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(keys, 0, and_=B - 1)
    perm_r = torch.empty_like(keys)
    off_ptr = L.ref_malloc(ref.CUDA, 4 * (4 * B + 1), 1)             # host-pinned, as the reference requires
    uniq = {}
    t_ref = time_ref(L, lambda: uniq.__setitem__("r", L.ref_block_mkperm(ref.CUDA, vp(keys.data_ptr()), n, n, B, vp(perm_r.data_ptr()), vp(off_ptr))))
    off_r = np.ctypeslib.as_array((ctypes.c_uint32 * (4 * B + 1)).from_address(off_ptr)).copy()
    out = {}
    t_ours = time_ours(lambda: out.__setitem__("o", dr.block_mkperm(keys, n, B)))
    torch.cuda.synchronize()
    perm_o, table = out["o"]
    perm_o = perm_o.view(torch.int32)
    tab_o = table.numpy().astype(np.uint32).reshape(-1, 4)
    tab_r = off_r[:4 * uniq["r"]].reshape(-1, 4)
    # {bucket, start, size} rows: the reference's CUDA kernel appends rows through an atomic counter
    # (resources/mkperm.cuh:309-317), so their order is arbitrary there; ours are in ascending bucket
    # order like the reference's CPU backend. Compare as sets of rows.
    tab_r = tab_r[np.argsort(tab_r[:, 0], kind="stable")]
    ok_table = (uniq["r"] == tab_o.shape[0] and np.array_equal(tab_o[:, :3], tab_r[:, :3])
                and np.all(np.diff(tab_o[:, 0].astype(np.int64)) > 0))
    # same elements per bucket: sorting each permutation by (key, index) must give the same array
    k64 = keys.to(torch.int64)
    pr, po = perm_r.to(torch.int64), perm_o.to(torch.int64)
    canon_r = torch.sort(k64[pr] * n + pr).values
    canon_o = torch.sort(k64[po] * n + po).values
    grouped_o = bool(torch.all(k64[po][1:] >= k64[po][:-1]))
    grouped_r = bool(torch.all(k64[pr][1:] >= k64[pr][:-1]))
    ok_sets = torch.equal(canon_r, canon_o)
    verdict = ("identical" if ok_table and ok_sets and grouped_o and grouped_r else
               f"MISMATCH (table {ok_table}, sets {ok_sets}, grouped ours {grouped_o} / reference {grouped_r}; "
               f"4th table word ours {tab_o[:2, 3].tolist()} reference {tab_r[:2, 3].tolist()})")
    report(f"block_mkperm {B} buckets", n, 12, t_ref, t_ours, f"{uniq['r']} buckets, table {{id,start,size}} + per-bucket contents " + verdict)
    L.ref_free(off_ptr)
    del keys, perm_r, canon_r, canon_o, k64, pr, po

    # ---- dr.argsort of 2^26 u32: the reference's _radix_sort (drjit/__init__.py:1698-1772) --------------
    # 4 x (digit kernel, jit_block_mkperm(256 buckets) of the reference's CUDA backend, one gather per
    # carried array); the digit extraction and the gathers are JIT kernels in the reference, torch
    # elementwise / index kernels stand in for them here (same memory traffic).
    n = 1 << (26 - S)
    keys = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(keys, 0)
    perm_r = torch.empty_like(keys)
    res = {}

    def ref_sort():
        o, idx = keys, torch.arange(n, dtype=torch.int32, device=dev)
        for shift in (0, 8, 16, 24):
            digit = ((o >> shift) & 255).contiguous()
            torch.cuda.synchronize()
            L.ref_block_mkperm(ref.CUDA, vp(digit.data_ptr()), n, n, 256, vp(perm_r.data_ptr()), None)
            L.ref_sync()
            pl = perm_r.long()
            o, idx = o[pl], idx[pl]
        res["r"] = (o, idx)

    t_ref = time_ref(L, ref_sort)
    t_ours = time_ours(lambda: res.__setitem__("o", ops.sort_with_indices(keys, vt=VarType.UInt32)))
    torch.cuda.synchronize()
    ok = torch.equal(res["r"][0], res["o"][0]) and torch.equal(res["r"][1], res["o"][1])
    report("argsort u32 (keys + index)", n, 80, t_ref, t_ours, "sorted keys and permutation " + ("bit-exact" if ok else "MISMATCH"))
    del keys, perm_r, res

    # ---- scatter-add histogram -------------------------------------------------------------------------
    n = 1 << (28 - S); bins = 1 << max(4, 20 - S)
    val = torch.empty(n, dtype=torch.float32, device=dev); ops.fill_fmix32(val, 1)
    idx = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(idx, 0, xor=0x85EBCA6B, and_=bins - 1)
    b_r = torch.zeros(bins, dtype=torch.float32, device=dev); b_o = torch.zeros_like(b_r)

    def run_ref():
        b_r.zero_(); torch.cuda.synchronize()
        rv = L.ref_scatter_reduce(ref.CUDA, VT["f32"], OP["add"], 0, vp(b_r.data_ptr()), bins, vp(val.data_ptr()), vp(idx.data_ptr()), n)
        if rv != 0:
            raise RuntimeError(L.ref_last_error().decode())
    try:
        t_ref = time_ref(L, run_ref)
        t_ours = time_ours(lambda: (b_o.zero_(), dr.scatter_add(b_o, val, idx)))
        rel = float(((b_o - b_r).abs() / b_r.abs().clamp(min=1)).max())
        report("scatter_add f32 -> 2^20 bins", n, 8, t_ref, t_ours, f"max rel diff {rel:.1e} (ref: JIT-compiled kernel, incl. zeroing)")
    except Exception as e:  # the tracer path needs the PTX JIT; report rather than fail
        print(f"scatter_add: reference tracer path unavailable here ({e})")


if __name__ == "__main__":
    main()
