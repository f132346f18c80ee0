#!/usr/bin/env python
"""A/B of the JIT scatter-reduce template: dr.scatter_reduce(Add / Max, u32,
ReduceMode.Local) traced and compiled by the reference's JIT, once with the reference's own template
(oracle/_ref) and once with the B200 template linked in (oracle/_ref_b200, seam_scatter_b200.cpp:
match.any + redux.sync). Kernel times come from JitFlag::KernelHistory (CUDA events around the JIT
kernel). One subprocess per library (same symbols).

    python scripts/scatter_template_ab.py
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BODY = r"""
import ctypes, sys
import numpy as np
from oracle import capi, ref
from oracle.capi import OP, VT
L = ref.lib(cuda=True, llvm=False)
vp = ctypes.c_void_p
CUDA, LOCAL = ref.CUDA, 2
n = 1 << 26
def dev(a):
    p = L.ref_malloc(CUDA, a.nbytes, 0); L.ref_memcpy(CUDA, vp(p), a.ctypes.data_as(vp), a.nbytes); return p
val = (capi.fmix32(n, xor=11) & 0xFFFF).astype(np.uint32); d_val = dev(val)
mask = (capi.fmix32(n, xor=3) & 3 != 0).astype(np.uint8); d_mask = dev(mask)          # 75 % active
N = 64
b = (ctypes.c_uint32 * N)(); t = (ctypes.c_uint32 * N)(); s = (ctypes.c_uint32 * N)(); ms = (ctypes.c_float * N)()
for bins in (1, 64, 1 << 12, 1 << 20):
    idx = (capi.fmix32(n, xor=0x85EBCA6B) % np.uint32(bins)).astype(np.uint32); d_idx = dev(idx)
    for op, masked in (("add", False), ("add", True), ("max", True)):
        d_t = dev(np.zeros(bins, np.uint32))
        best = 1e9
        for rep in range(4):
            L.ref_set_flag(1 << 15, 1)
            assert L.ref_scatter_reduce_masked(CUDA, VT["u32"], OP[op], LOCAL, vp(d_t), bins, vp(d_val), vp(d_idx),
                                               vp(d_mask) if masked else None, n) == 0
            L.ref_sync()
            cnt = L.ref_kernel_history(b, t, s, ms, N)
            jit = [ms[i] for i in range(cnt) if t[i] == 0 and s[i] == n]
            L.ref_set_flag(1 << 15, 0)
            if rep and jit: best = min(best, jit[-1])
        print(f"{sys.argv[1]:10s} u32 {op:3s} {'masked 75%' if masked else 'unmasked  '} 2^26 values -> {bins:8d} bins  {best:8.3f} ms  {n / best / 1e6:8.1f} Gelem/s", flush=True)
        L.ref_free(vp(d_t))
    L.ref_free(vp(d_idx))
"""

for name, d in (("reference", "_ref"), ("b200", "_ref_b200")):
    env = dict(os.environ, ORACLE_REF_DIR=os.path.join(ROOT, "oracle", d), PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", BODY, name], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    print(out.stdout, end="")
    if out.returncode:
        print(out.stderr[-2000:])
