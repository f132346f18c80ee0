"""The drop-in boundary proven in situ: oracle/_ref_b200/libdrjit-core.so is the reference drjit-core
with the eight primitive methods of CUDAThreadState (src/cuda_ts.cpp:129-1026) bound to
libdrjit_b200.so exactly as INTEGRATION.md shows (oracle/ref_build/seam_b200.cpp; everything else
is the unmodified reference). On the GPU box

 * the reference's OWN test programs tests/reductions.cpp (:109-406: block sums, prefix sums, compress,
   mkperm on the fmix32 size grid) and tests/vcall.cpp (jit_var_call_reduce -> block_mkperm, poke /
   aggregate parameter blocks) run their CUDA cases through the new kernels, and
 * jit_block_reduce / jit_block_prefix_reduce / jit_compress / jit_block_mkperm(JitBackend::CUDA) of
   that library are checked against the C oracle, with JitFlag::KernelHistory entries carrying the
   reference's KernelType tags and JitFlag::LaunchBlocking honoured (cuda_ts.cpp:19-46), and
 * dr.scatter_reduce traced by the reference's JIT with the B200 scatter template linked in
   (oracle/ref_build/seam_scatter_b200.cpp, SURVEY section 8 row f3: match.any + redux.sync for
   ReduceMode::Local on 32-bit integers; tests/mem.cpp covers the forwarded float paths) matches the
   oracle for every (type, op) pair, masked and unmasked, and the generated PTX contains redux.sync.

The second half runs in a subprocess: the patched and the unmodified libdrjit-core.so cannot share
one process (same symbols)."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200 = os.path.join(ROOT, "oracle", "_ref_b200")


def _need(path):
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, ROOT)} not built (make -C oracle/ref_build b200)")
    return path


@pytest.mark.parametrize("name", ["reductions", "vcall", "mem"])
def test_reference_test_program_through_b200_kernels(name, tmp_path):
    exe = _need(os.path.join(B200, f"test_{name}"))
    (tmp_path / f"out_{name}").mkdir()          # the harness writes its logs to out_<name>/ of the cwd
    env = dict(os.environ, DRJIT_B200_INSITU_TRACE="1")
    out = subprocess.run([exe, "-c"], cwd=tmp_path, capture_output=True, text=True, timeout=900, env=env)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    m = re.search(r"Passed (\d+)/(\d+) tests", out.stdout)
    assert m, tail
    assert int(m.group(1)) == int(m.group(2)) and int(m.group(1)) > 0, tail      # CUDA cases ran and all passed
    assert "FAILED" not in out.stdout


BODY = r"""
import ctypes, sys
import numpy as np
from oracle import capi, ref
from oracle.capi import OP, VT

assert ref.REF_DIR.endswith("_ref_b200")
L = ref.lib(cuda=True, llvm=False)
assert ref.has_backend(ref.CUDA), "reference CUDA backend did not initialise"
vp = ctypes.c_void_p
CUDA = ref.CUDA

# the primitives of this process must be libdrjit_b200.so's
maps = open("/proc/self/maps").read()
assert "libdrjit_b200.so" in maps and "_ref_b200/libdrjit-core.so" in maps

def dev(a):
    p = L.ref_malloc(CUDA, max(a.nbytes, 4), 0)
    L.ref_memcpy(CUDA, vp(p), a.ctypes.data_as(vp), a.nbytes)
    return p

def host(p, n, dt):
    a = np.empty(n, dt)
    L.ref_sync()
    L.ref_memcpy(CUDA, a.ctypes.data_as(vp), vp(p), a.nbytes)
    return a

KH, LB = 1 << 15, 1 << 16
L.ref_set_flag(KH, 1)
L.ref_kernel_history_clear()
expect_types = []

for n in (1, 1000, (1 << 20) + 77):
    u = capi.fmix32(n)
    d_in = dev(u); d_out = L.ref_malloc(CUDA, 4 * n, 0)
    # jit_block_reduce: full reduction and blocks of 256 / 3
    for bs in sorted({n, min(n, 256), min(n, 3)}):
        assert L.ref_block_reduce(CUDA, VT["u32"], OP["add"], n, bs, vp(d_in), vp(d_out)) == 0, L.ref_last_error()
        exp = capi.block_reduce("u32", "add", u, bs)
        assert np.array_equal(host(d_out, exp.size, np.uint32), exp), ("block_reduce", n, bs)
        if bs > 1:
            expect_types.append(1)      # KernelType::BlockReduce
    # jit_block_prefix_reduce (size, block_size) order; exclusive / inclusive / reverse; segmented
    for bs in sorted({n, min(n, 1000)}):
        for ex, rev in ((1, 0), (0, 0), (1, 1)):
            assert L.ref_block_prefix_reduce(CUDA, VT["u32"], OP["add"], n, bs, ex, rev, vp(d_in), vp(d_out)) == 0
            exp = capi.block_prefix_reduce("u32", "add", u, bs, bool(ex), bool(rev))
            assert np.array_equal(host(d_out, n, np.uint32), exp), ("prefix", n, bs, ex, rev)
            if bs > 1:
                expect_types.append(2)  # KernelType::BlockPrefixReduce
    # jit_compress
    m = capi.mask_u8(n, 77)
    d_m = dev(np.concatenate([m, np.zeros(4096, np.uint8)]))
    cnt = L.ref_compress(CUDA, vp(d_m), n, vp(d_out))
    exp = capi.compress(m)
    assert cnt == exp.size and np.array_equal(host(d_out, cnt, np.uint32), exp), ("compress", n)
    expect_types.append(5)              # KernelType::Compress
    # jit_block_mkperm: stable below 1816 buckets -> bit-exact against the oracle (= LLVM backend)
    for B in (37, 4096):
        keys = capi.fmix32(n) % np.uint32(B)
        d_k = dev(keys)
        d_off = L.ref_malloc(CUDA, 4 * (4 * B + 1), 1)      # shared = pinned host memory
        unique = L.ref_block_mkperm(CUDA, vp(d_k), n, n, B, vp(d_out), vp(d_off))
        eperm, eoff, eunique = capi.block_mkperm(keys, n, B)
        assert unique == eunique, ("mkperm unique", n, B)
        off = np.ctypeslib.as_array(ctypes.cast(d_off, ctypes.POINTER(ctypes.c_uint32)), shape=(4 * B + 1,)).copy()
        assert np.array_equal(off[:4 * unique], eoff[:4 * unique]) and off[4 * B] == unique
        p = host(d_out, n, np.uint32)
        if B * 4 * 32 <= 227 * 1024 or n < (1 << 18):
            assert np.array_equal(p, eperm), ("mkperm perm", n, B)
        else:
            assert np.array_equal(np.sort(p), np.arange(n, dtype=np.uint32)) and np.all(np.diff(keys[p].astype(np.int64)) >= 0)
        expect_types.append(6)          # KernelType::MkPerm
        L.ref_free(vp(d_k)); L.ref_free(vp(d_off))
    L.ref_free(vp(d_in)); L.ref_free(vp(d_out)); L.ref_free(vp(d_m))

# dot product through the variable-level entry (jit_var_reduce_dot)
f = capi.unit_f32(100003)
d_f = dev(f)
out = np.zeros(1, np.float32)
assert L.ref_reduce_dot(CUDA, VT["f32"], vp(d_f), vp(d_f), f.size, out.ctypes.data_as(vp)) == 0
exp = float(np.dot(f.astype(np.float64), f.astype(np.float64)))
assert abs(float(out[0]) - exp) <= 1e-6 * 17 * exp
expect_types.append(3)                  # KernelType::Dot

# ---- dr.kernel_history(): one entry per primitive call, reference KernelType tags, timed
N = 4096
b = (ctypes.c_uint32 * N)(); t = (ctypes.c_uint32 * N)(); s = (ctypes.c_uint32 * N)(); ms = (ctypes.c_float * N)()
L.ref_sync()
cnt = L.ref_kernel_history(b, t, s, ms, N)
types = [t[i] for i in range(cnt)]
prim = [x for x in types if x in (1, 2, 3, 5, 6)]
assert prim == expect_types, (prim, expect_types)
assert all(b[i] == CUDA for i in range(cnt)) and all(ms[i] >= 0 for i in range(cnt))
assert any(ms[i] > 0 for i in range(cnt))

# ---- LaunchBlocking: the call returns with the result already in memory (no explicit sync)
L.ref_set_flag(KH, 0); L.ref_set_flag(LB, 1)
n = 1 << 22
u = capi.fmix32(n); d_in = dev(u)
d_pin = L.ref_malloc(CUDA, 4, 1)
assert L.ref_block_reduce(CUDA, VT["u32"], OP["add"], n, n, vp(d_in), vp(d_pin)) == 0
got = ctypes.cast(d_pin, ctypes.POINTER(ctypes.c_uint32))[0]
assert got == int(u.sum(dtype=np.uint32)), "LaunchBlocking: result not ready on return"
L.ref_set_flag(LB, 0)
print("insitu ok:", cnt, "history entries,", len(expect_types), "primitive calls")
"""


SCATTER_BODY = r"""
import ctypes
import numpy as np
from oracle import capi, ref
from oracle.capi import OP, VT

assert ref.REF_DIR.endswith("_ref_b200")
L = ref.lib(cuda=True, llvm=False)
assert ref.has_backend(ref.CUDA)
vp = ctypes.c_void_p
CUDA = ref.CUDA
KH = 1 << 15
LOCAL, DIRECT = 2, 1

def dev(a):
    p = L.ref_malloc(CUDA, max(a.nbytes, 4), 0)
    L.ref_memcpy(CUDA, vp(p), a.ctypes.data_as(vp), a.nbytes)
    return p

def host(p, n, dt):
    a = np.empty(n, dt)
    L.ref_sync()
    L.ref_memcpy(CUDA, a.ctypes.data_as(vp), vp(p), a.nbytes)
    return a

n = 100_003
checked = 0
for bins, pattern in ((1, "one counter"), (7, "few bins"), (1 << 12, "histogram")):
    idx = (capi.fmix32(n, xor=0x85EBCA6B) % np.uint32(bins)).astype(np.uint32)
    mask = (capi.fmix32(n, xor=3) & 3 != 0).astype(np.uint8)          # 75 % active
    d_idx, d_mask = dev(idx), dev(mask)
    for vt, dt in (("u32", np.uint32), ("i32", np.int32)):
        raw = capi.fmix32(n, xor=11)
        val = (raw & 0xFFFF).astype(np.uint32).view(dt) if vt == "u32" else ((raw & 0xFFFF).astype(np.int64) - 0x8000).astype(np.int32)
        d_val = dev(val)
        for op in ("add", "min", "max", "and", "or"):
            for use_mask in (False, True):
                init = {"add": 5, "min": 0x7fff0000 if vt == "i32" else 0xffff0000, "max": -0x7fff0000 if vt == "i32" else 0,
                        "and": -1 if vt == "i32" else 0xffffffff, "or": 0}[op]
                target = np.full(bins, init, dtype=np.int64).astype(dt)
                exp = capi.scatter_reduce(vt, op, target, val, idx, mask if use_mask else None)
                d_t = dev(target)
                L.ref_set_flag(KH, 1)
                rv = L.ref_scatter_reduce_masked(CUDA, VT[vt], OP[op], LOCAL, vp(d_t), bins, vp(d_val), vp(d_idx),
                                                 vp(d_mask) if use_mask else None, n)
                assert rv == 0, L.ref_last_error()
                got = host(d_t, bins, dt)
                assert np.array_equal(got, exp), (pattern, vt, op, use_mask)
                # the JIT kernel of this call was generated by the B200 template
                assert L.ref_kernel_history_ir_count(b"redux.sync") >= 1, (pattern, vt, op, use_mask)
                L.ref_set_flag(KH, 0)
                L.ref_free(vp(d_t))
                checked += 1
        L.ref_free(vp(d_val))
    # forwarded paths stay the reference's: f32 add in Local and Direct mode, u32 add in Direct mode
    f = capi.unit_f32(n); d_f = dev(f)
    for mode in (LOCAL, DIRECT):
        d_t = dev(np.zeros(bins, np.float32))
        L.ref_set_flag(KH, 1)
        assert L.ref_scatter_reduce_masked(CUDA, VT["f32"], OP["add"], mode, vp(d_t), bins, vp(d_f), vp(d_idx), None, n) == 0
        got = host(d_t, bins, np.float32)
        exp = capi.scatter_reduce("f32", "add", np.zeros(bins, np.float32), f, idx, acc64=True)
        assert np.all(np.abs(got - exp) <= 1e-4 * np.maximum(np.abs(exp), 1)), (pattern, mode)
        assert L.ref_kernel_history_ir_count(b"redux.sync") == 0
        L.ref_set_flag(KH, 0)
        L.ref_free(vp(d_t))
    L.ref_free(vp(d_f)); L.ref_free(vp(d_idx)); L.ref_free(vp(d_mask))
print("scatter template ok:", checked, "integer cases")
"""


def _run_body(body, marker):
    _need(os.path.join(B200, "libref_shim.so"))
    env = dict(os.environ, ORACLE_REF_DIR=B200, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-c", body], cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert marker in out.stdout


def test_jit_scatter_template_redux_path_vs_oracle():
    _run_body(SCATTER_BODY, "scatter template ok")


def test_jit_entry_points_of_patched_library_vs_oracle():
    _need(os.path.join(B200, "libref_shim.so"))
    env = dict(os.environ, ORACLE_REF_DIR=B200, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-c", BODY], cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "insitu ok" in out.stdout
