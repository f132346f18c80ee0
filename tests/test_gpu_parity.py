"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Integer / index results: bit-exact. Floating point: relative tolerance 1e-6 * log2(N) for f32
(north_star), 2e-3-scale for f16, 1e-13-scale for f64, stated next to each check.
The grids follow the reference's own tests (ext/drjit-core/tests/reductions.cpp,
tests/test_reduction.py, tests/test_memop.py)."""
import ctypes

import numpy as np
import pytest
import torch

import drjit_b200 as dr
from drjit_b200 import ReduceMode, ReduceOp, VarType
from oracle import capi
from tests.gpu_util import NP, OPS, VT, to_dev, to_np
from tests.golden import loader
from tests.golden.make_golden import digest, make_input

pytestmark = pytest.mark.gpu

# ext/drjit-core/tests/reductions.cpp:73-76
RED_SIZES = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 32, 60, 128, 250, 333, 1024, 16384,
             16388 * 10, 9973 * 17, 98973 * 17 * 3]
INT_TYPES = ["u8", "i32", "u32", "i64", "u64"]
FLT_TYPES = ["f16", "f32", "f64"]


def ftol(vt, n):
    base = {"f16": 1.5e-3, "f32": 1e-6, "f64": 1e-14}[vt]
    return base * max(1.0, np.log2(max(n, 2)))


def assert_close(got, exp, vt, n, what):
    got = got.astype(np.float64); exp = exp.astype(np.float64)
    both_inf = np.isinf(got) & (got == exp)     # e.g. f16 sums beyond 65504 overflow identically
    got = np.where(both_inf, 0.0, got); exp = np.where(both_inf, 0.0, exp)
    scale = np.maximum(np.abs(exp), 1.0)
    err = np.abs(got - exp) / scale
    assert np.all(err <= ftol(vt, n)), f"{what}: max rel err {err.max():.3e} > {ftol(vt, n):.3e}"


# --------------------------------------------------------------------------- block_reduce
@pytest.mark.parametrize("vt", ["u32", "u64"])
def test_block_sum_reference_grid(vt):
    """reductions.cpp:122-151: every (size, block_size) pair of red_sizes, exact"""
    for size in RED_SIZES:
        x = capi.fmix32(size) if vt == "u32" else capi.fmix32_u64(size)
        xd = to_dev(x, vt)
        for bs in RED_SIZES:
            if bs > size:
                continue
            got = to_np(dr.block_reduce(ReduceOp.Add, xd, bs, vt=VT[vt]), vt)
            assert np.array_equal(got, capi.block_reduce(vt, "add", x, bs)), (size, bs)


@pytest.mark.parametrize("vt", INT_TYPES)
@pytest.mark.parametrize("op", ["add", "mul", "min", "max", "and", "or"])
def test_block_reduce_int_ops(vt, op):
    for size in [1, 7, 333, 16384 + 4, 9973 * 17]:
        x = make_input(vt, size)
        for misalign in (0, 1):
            xd = to_dev(x, vt, misalign)
            for bs in [1, 2, 3, 7, 32, 250, 256, 1024, 5000, 65536, size]:
                if bs > size:
                    continue
                got = to_np(dr.block_reduce(OPS[op], xd, bs, vt=VT[vt]), vt)
                assert np.array_equal(got, capi.block_reduce(vt, op, x, bs)), (size, bs, misalign)


@pytest.mark.parametrize("vt", FLT_TYPES)
@pytest.mark.parametrize("op", ["add", "mul", "min", "max"])
def test_block_reduce_float_ops(vt, op):
    for size in [1, 7, 333, 16384 + 4, 9973 * 17]:
        x = make_input(vt, size)
        if op == "mul":
            x = (1.0 + (x.astype(np.float64) - 0.5) * 1e-3).astype(NP[vt])  # keep products finite
        xd = to_dev(x, vt)
        for bs in [2, 3, 32, 250, 256, 1024, 5000, size]:
            if bs > size:
                continue
            got = to_np(dr.block_reduce(OPS[op], xd, bs, vt=VT[vt]), vt)
            exp = capi.block_reduce(vt, op, x, bs, acc64=True)
            if op in ("min", "max"):
                assert np.array_equal(got, exp), (size, bs)
            elif op == "mul":
                # a product of n factors accumulates n roundings (each 2^-p relative): tolerance
                # eps * sqrt(n) * 4 on top of the sum tolerance (the oracle multiplies in f64)
                eps = {"f16": 2.0 ** -11, "f32": 2.0 ** -24, "f64": 2.0 ** -53}[vt]
                rel = np.abs(got.astype(np.float64) - exp.astype(np.float64)) / np.maximum(np.abs(exp.astype(np.float64)), 1e-30)
                tol = ftol(vt, bs) + 4 * eps * np.sqrt(bs)
                assert np.all(rel <= tol), f"{vt} mul size={size} bs={bs}: {rel.max():.3e} > {tol:.3e}"
            else:
                assert_close(got, exp, vt, bs, f"{vt} {op} size={size} bs={bs}")


@pytest.mark.parametrize("vt", ["u8", "f16", "i32", "u32", "f32", "u64", "f64"])
def test_block_reduce_tiny_blocks(vt):
    """block sizes that fill a 16-byte vector or an integer fraction of one (dedicated kernel):
    every op of the type, whole-vector and ragged sizes, aligned and offset input"""
    ops_for = ["add", "mul", "min", "max"] + ([] if vt[0] == "f" else ["and", "or"])
    for size in [64, 4096 * 3, (1 << 20) + 16, (1 << 20) + 5]:
        x = make_input(vt, size)
        for misalign in (0, 1):
            xd = to_dev(x, vt, misalign)
            for bs in (2, 3, 4, 5, 6, 7, 8, 16):      # (3, 5, 6, 7: the short-block kernel + a general-path tail)
                for op in ops_for:
                    got = to_np(dr.block_reduce(OPS[op], xd, bs, vt=VT[vt]), vt)
                    exp = capi.block_reduce(vt, op, x, bs, acc64=(vt[0] == "f"))
                    if vt[0] == "f" and op in ("add", "mul"):
                        assert_close(got, exp, vt, bs, f"{vt} {op} bs={bs} size={size}")
                    else:
                        assert np.array_equal(got, exp), (vt, op, bs, size, misalign)


def test_block_reduce_golden_fixture():
    """against outputs of the unmodified reference (tests/golden/ref_llvm.npz), integer digests"""
    digests, _ = loader.load()
    for vt in ["u8", "i32", "u32", "i64", "u64"]:
        for n in [333, 5000, 16384 + 4]:
            x = make_input(vt, n); xd = to_dev(x, vt)
            for bs in [2, 7, 60, 250, 1024]:
                if bs > n:
                    continue
                for op in ["add", "mul", "min", "max", "and", "or"]:
                    got = to_np(dr.block_reduce(OPS[op], xd, bs, vt=VT[vt]), vt)
                    assert digest(got)[0] == digests[f"br/{vt}/{op}/{n}/{bs}"], (vt, n, bs, op)
                    for ex in (0, 1):
                        for rev in (0, 1):
                            got = to_np(dr.block_prefix_reduce(OPS[op], xd, bs, ex, rev, vt=VT[vt]), vt)
                            assert digest(got)[0] == digests[f"bp/{vt}/{op}/{n}/{bs}/{ex}{rev}"], (vt, n, bs, op, ex, rev)


def test_block_reduce_errors_and_edges():
    """cuda_ts.cpp:200-213: empty no-op, invalid block size raises, block_size == 1 copies"""
    x = to_dev(capi.fmix32(10), "u32")
    assert dr.block_reduce(ReduceOp.Add, x[:0], 1).numel() == 0
    for bs in (0, 11):
        with pytest.raises(RuntimeError, match="invalid block size"):
            dr.block_reduce(ReduceOp.Add, x, bs, vt=VarType.UInt32)
        with pytest.raises(RuntimeError, match="invalid block size"):
            dr.block_prefix_reduce(ReduceOp.Add, x, bs, vt=VarType.UInt32)
    assert torch.equal(dr.block_reduce(ReduceOp.Max, x, 1, vt=VarType.UInt32), x)
    f = torch.ones(8, device="cuda")
    with pytest.raises(RuntimeError, match="no existing kernel"):
        dr.block_reduce(ReduceOp.And, f, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dr.sum(torch.ones(4))


def test_full_reduction_large():
    """dr.sum at 2^28 f32 (BASELINE config 1): tolerance 1e-6*log2(N); u32: exact"""
    n = 1 << 28
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    dr.ops.fill_fmix32(x, 1)
    got = float(dr.sum(x).item())
    exp = float(capi.block_reduce("f32", "add", capi.unit_f32(n), n, acc64=True)[0])
    assert abs(got - exp) <= 1e-6 * 28 * abs(exp), (got, exp)
    # block_reduce(Add, 256): every block against float64 torch reduction
    out = dr.block_reduce(ReduceOp.Add, x, 256)
    ref = x.view(-1, 256).to(torch.float64).sum(dim=1)
    assert torch.all((out.to(torch.float64) - ref).abs() <= 1e-6 * 8 * ref.abs().clamp_min(1.0))
    # u32, exact (wraps mod 2^32)
    xi = torch.empty(n, dtype=torch.int32, device="cuda")
    dr.ops.fill_fmix32(xi, 0)
    got = int(dr.sum(xi, vt=VarType.UInt32).item()) & 0xFFFFFFFF
    assert got == int(capi.fmix32(n).sum(dtype=np.uint32))
    assert (int(dr.min(xi, vt=VarType.UInt32).item()) & 0xFFFFFFFF) == int(capi.fmix32(n).min())
    assert (int(dr.max(xi, vt=VarType.Int32).item())) == int(capi.fmix32(n).view(np.int32).max())


def test_all_any():
    """reductions.cpp:78-107: single-bit flips over sizes 23 i^3 + 1"""
    rng = np.random.default_rng(1)
    for i in list(range(0, 100, 9)) + [99]:
        size = 23 * i ** 3 + 1
        f = torch.zeros(size, dtype=torch.bool, device="cuda")
        t = torch.ones(size, dtype=torch.bool, device="cuda")
        assert dr.all(t) and dr.any(t) and not dr.all(f) and not dr.any(f)
        for _ in range(4):
            k = int(rng.integers(0, size))
            f[k] = True; t[k] = False
            if size == 1:
                assert not dr.all(t) and not dr.any(t) and dr.all(f) and dr.any(f)
            else:
                assert not dr.all(t) and dr.any(t) and not dr.all(f) and dr.any(f)
            f[k] = False; t[k] = True
    # misaligned mask (sub-range of an allocation)
    t = torch.ones(1000, dtype=torch.bool, device="cuda")[3:]
    assert dr.all(t)
    t[500] = False
    assert not dr.all(t) and dr.any(t)


@pytest.mark.parametrize("vt", FLT_TYPES)
def test_dot(vt):
    for n in [1, 5, 100, 5000, 1 << 20, (1 << 22) + 3]:
        a = make_input(vt, n); b = make_input(vt, n)[::-1].copy()
        for misalign in (0, 1):
            got = to_np(dr.dot(to_dev(a, vt, misalign), to_dev(b, vt, misalign)), vt)[0]
            exp = capi.reduce_dot(vt, a, b, acc64=True)
            tol = ftol(vt, n) * max(1.0, abs(float(exp)))
            if vt == "f16":
                tol = max(tol, 2e-3 * abs(float(exp)))
            if np.isinf(float(exp)) or np.isinf(float(got)):   # f16 result beyond 65504
                assert float(got) == float(exp), (vt, n, got, exp)
                continue
            assert abs(float(got) - float(exp)) <= tol, (vt, n, got, exp)
    got = to_np(dr.dot(to_dev(make_input("f32", 1001), "f32"), to_dev(make_input("f32", 1001), "f32", 1)), "f32")
    assert abs(float(got[0]) - float(capi.reduce_dot("f32", make_input("f32", 1001), make_input("f32", 1001), acc64=True))) < 1e-3


# --------------------------------------------------------------------------- prefix reductions
@pytest.mark.parametrize("vt", ["u32", "u64"])
@pytest.mark.parametrize("exclusive", [0, 1])
@pytest.mark.parametrize("reverse", [0, 1])
def test_block_prefix_sum_reference_grid(vt, exclusive, reverse):
    """reductions.cpp:153-267: all four variants over the red_sizes grid, exact"""
    sizes = [s for s in RED_SIZES if s <= 9973 * 17]
    for size in sizes:
        x = capi.fmix32(size) if vt == "u32" else capi.fmix32_u64(size)
        xd = to_dev(x, vt)
        for bs in sizes:
            if bs > size:
                continue
            got = to_np(dr.block_prefix_reduce(ReduceOp.Add, xd, bs, exclusive, reverse, vt=VT[vt]), vt)
            assert np.array_equal(got, capi.block_prefix_reduce(vt, "add", x, bs, exclusive, reverse)), (size, bs)


def test_prefix_sum_largest_reference_size_and_inplace():
    n = RED_SIZES[-1]
    x = capi.fmix32(n)
    for misalign in (0, 1):
        xd = to_dev(x, "u32", misalign)
        for ex in (0, 1):
            for rev in (0, 1):
                for bs in (n, 16388 * 10, 1000):
                    got = to_np(dr.block_prefix_reduce(ReduceOp.Add, xd, bs, ex, rev, vt=VarType.UInt32), "u32")
                    assert np.array_equal(got, capi.block_prefix_reduce("u32", "add", x, bs, ex, rev)), (misalign, ex, rev, bs)
    # in place (jit.h:2346-2364)
    xd = to_dev(x, "u32")
    dr.block_prefix_reduce(ReduceOp.Add, xd, n, True, False, vt=VarType.UInt32, out=xd)
    assert np.array_equal(to_np(xd, "u32"), capi.block_prefix_reduce("u32", "add", x, n, 1, 0))


@pytest.mark.parametrize("vt", INT_TYPES)
@pytest.mark.parametrize("op", ["add", "mul", "min", "max", "and", "or"])
def test_block_prefix_reduce_int_ops(vt, op):
    for size in [1, 333, 40000 + 3]:
        x = make_input(vt, size); xd = to_dev(x, vt)
        for bs in [1, 2, 3, 7, 250, 256, 4096, 10000, size]:
            if bs > size:
                continue
            for ex, rev in ((0, 0), (1, 0), (0, 1), (1, 1)):
                got = to_np(dr.block_prefix_reduce(OPS[op], xd, bs, ex, rev, vt=VT[vt]), vt)
                assert np.array_equal(got, capi.block_prefix_reduce(vt, op, x, bs, ex, rev)), (size, bs, ex, rev)


@pytest.mark.parametrize("vt", FLT_TYPES)
@pytest.mark.parametrize("op", ["add", "min", "max"])
def test_block_prefix_reduce_float_ops(vt, op):
    for size in [333, 40000 + 3]:
        x = make_input(vt, size); xd = to_dev(x, vt)
        for bs in [2, 7, 256, 5000, size]:
            if bs > size:
                continue
            for ex, rev in ((0, 0), (1, 1)):
                got = to_np(dr.block_prefix_reduce(OPS[op], xd, bs, ex, rev, vt=VT[vt]), vt)
                exp = capi.block_prefix_reduce(vt, op, x, bs, ex, rev, acc64=True)
                if op == "add":
                    assert_close(got, exp, vt, bs, f"{vt} prefix size={size} bs={bs}")
                else:
                    assert np.array_equal(got, exp), (size, bs, ex, rev)


def test_prefix_literals():
    """tests/test_memop.py:734-756, tests/test_reduction.py:370-387"""
    x = torch.tensor([1, 2, 3, 4, 5, 6], dtype=torch.float32, device="cuda")
    exp_sum = {1: [1, 2, 3, 4, 5, 6], 2: [3, 7, 11], 3: [6, 15], 4: [10, 11], 5: [15, 6], 6: [21]}
    exp_psum = {1: [0] * 6, 2: [0, 1, 0, 3, 0, 5], 3: [0, 1, 3, 0, 4, 9], 4: [0, 1, 3, 6, 0, 5],
                5: [0, 1, 3, 6, 10, 0], 6: [0, 1, 3, 6, 10, 15]}
    for dtype in (torch.float32, torch.float64, torch.int32, torch.int64, torch.float16):
        xx = x.to(dtype)
        for bs in range(1, 7):
            assert dr.block_sum(xx, bs).tolist() == exp_sum[bs]
            assert dr.block_prefix_sum(xx, bs).tolist() == exp_psum[bs]
    assert dr.prefix_sum(x[:3]).tolist() == [0, 1, 3]
    assert dr.cumsum(x[:3]).tolist() == [1, 3, 6]


def test_prefix_sum_2_30_properties():
    """exclusive u32 prefix sum at the BASELINE size 2^30 through size-independent properties:
    out[0] == 0, out[i+1] - out[i] == in[i] (mod 2^32), out[n-1] + in[n-1] == sum"""
    n = 1 << 30
    x = torch.empty(n, dtype=torch.int32, device="cuda")
    dr.ops.fill_fmix32(x, 0)
    out = dr.prefix_sum(x, vt=VarType.UInt32)
    assert int(out[0].item()) == 0
    assert bool(torch.all(out[1:] - out[:-1] == x[:-1]))
    total = dr.sum(x, vt=VarType.UInt32)
    assert int((out[-1] + x[-1]).item()) == int(total.item())
    del out
    # first 2^22 entries bit-exact against the oracle
    exp = capi.block_prefix_reduce("u32", "add", capi.fmix32(1 << 22), 1 << 22, 1, 0)
    out = dr.prefix_sum(x[:1 << 22], vt=VarType.UInt32)
    assert np.array_equal(to_np(out, "u32"), exp)


@pytest.mark.parametrize("vt", ["u8", "f16", "i32", "u32", "f32", "i64", "f64"])
def test_block_prefix_reduce_short_blocks(vt):
    """block sizes 2..8 on large arrays (scanned inside one thread by a dedicated kernel, with a
    general-path tail): every op, all four variants, ragged sizes, in place"""
    ops_for = ["add", "min", "max"] + ([] if vt[0] == "f" else ["mul", "and", "or"])
    for size in [(1 << 17) + 13, (1 << 17) + 16 * 35]:
        x = make_input(vt, size); xd = to_dev(x, vt)
        for bs in range(2, 9):
            for op in ops_for:
                for ex, rev in ((0, 0), (1, 0), (0, 1), (1, 1)):
                    got = to_np(dr.block_prefix_reduce(OPS[op], xd, bs, ex, rev, vt=VT[vt]), vt)
                    exp = capi.block_prefix_reduce(vt, op, x, bs, ex, rev, acc64=(vt[0] == "f"))
                    if vt[0] == "f" and op == "add":
                        assert_close(got, exp, vt, bs, f"{vt} prefix bs={bs} size={size}")
                    else:
                        assert np.array_equal(got, exp), (vt, op, bs, ex, rev, size)
    y = xd.clone()
    dr.block_prefix_reduce(OPS["add"], y, 4, True, False, vt=VT[vt], out=y)
    exp = capi.block_prefix_reduce(vt, "add", x, 4, True, False, acc64=(vt[0] == "f"))
    if vt[0] == "f":
        assert_close(to_np(y, vt), exp, vt, 4, "in place")
    else:
        assert np.array_equal(to_np(y, vt), exp)


@pytest.mark.parametrize("vt", ["u8", "f16", "i32", "u32", "f32", "u64", "f64"])
def test_block_prefix_reduce_medium_blocks(vt):
    """blocks of 9 .. 32 Ki elements on arrays large enough for the group-per-block kernel (a group of
    2..32 lanes walks one block; blocks start inside 16-byte units, the last block is ragged): all four
    variants, several ops, in place; blocks past the kernel's limit fall through to the general path."""
    size = 3_000_017 if vt != "u8" else 6_000_029
    x = make_input(vt, size); xd = to_dev(x, vt)
    ops_for = ["add", "max"] + ([] if vt[0] == "f" else ["and"])
    for bs in [9, 17, 33, 100, 1000, 4099, 32768, 70001]:
        for op in ops_for:
            for ex, rev in ((0, 0), (1, 0), (0, 1), (1, 1)):
                got = to_np(dr.block_prefix_reduce(OPS[op], xd, bs, ex, rev, vt=VT[vt]), vt)
                exp = capi.block_prefix_reduce(vt, op, x, bs, ex, rev, acc64=(vt[0] == "f"))
                if vt[0] == "f" and op == "add":
                    assert_close(got, exp, vt, bs, f"{vt} prefix bs={bs} size={size}")
                else:
                    assert np.array_equal(got, exp), (vt, op, bs, ex, rev)
    for bs, ex, rev in ((1000, True, False), (37, False, True)):
        y = xd.clone()
        dr.block_prefix_reduce(OPS["max"], y, bs, ex, rev, vt=VT[vt], out=y)
        assert np.array_equal(to_np(y, vt), capi.block_prefix_reduce(vt, "max", x, bs, ex, rev)), ("in place", bs)


@pytest.mark.parametrize("vt", ["u8", "f16", "u32", "i32", "f32", "u64", "f64"])
def test_prefix_small_arrays_and_carry_form(vt):
    """Arrays of up to 128 KiB take one 1024-thread CTA with all loads in flight (prefix_small.cu), larger
    ones the tile kernel: sizes on both sides of every row / size limit, all four variants, in place, and
    the carry form of sharded scans (carry_in / total_out) on both paths."""
    item = np.dtype({"u8": np.uint8, "f16": np.float16, "u32": np.uint32, "i32": np.int32, "f32": np.float32,
                     "u64": np.uint64, "f64": np.float64}[vt]).itemsize
    V = 16 // item
    sizes = [1, V - 1 if V > 1 else 1, V, V + 1, 1024 * V - 1, 1024 * V, 1024 * V + 1, 2 * 1024 * V + 5, 4 * 1024 * V,
             8 * 1024 * V - 3, 8 * 1024 * V, 8 * 1024 * V + 1, 20 * 1024 * V + 7]
    ops_for = ["add", "max"] + ([] if vt[0] == "f" else ["or"])
    for size in sorted(set(sizes)):
        x = make_input(vt, size); xd = to_dev(x, vt)
        for op in ops_for:
            for ex, rev in ((1, 0), (0, 0), (1, 1), (0, 1)):
                got = to_np(dr.block_prefix_reduce(OPS[op], xd, size, ex, rev, vt=VT[vt]), vt)
                exp = capi.block_prefix_reduce(vt, op, x, size, ex, rev, acc64=(vt[0] == "f"))
                if vt[0] == "f" and op == "add":
                    assert_close(got, exp, vt, size, f"{vt} prefix size={size}")
                else:
                    assert np.array_equal(got, exp), (vt, op, size, ex, rev)
        y = xd.clone()
        dr.block_prefix_reduce(OPS["max"], y, size, True, False, vt=VT[vt], out=y)
        assert np.array_equal(to_np(y, vt), capi.block_prefix_reduce(vt, "max", x, size, 1, 0)), ("in place", size)
    if vt in ("u32", "u64", "i32"):
        # carry form: two pieces scanned one after the other == the scan of the whole (exact for integers)
        for n, cut in ((5000, 1232), (5000, 1234), (3 * 8 * 1024 * V, 8 * 1024 * V - 4), (100_000 + 3, 70_001)):
            x = make_input(vt, n); xd = to_dev(x, vt)
            exp = capi.block_prefix_reduce(vt, "add", x, n, 1, 0)
            out = torch.empty_like(xd)
            carry = torch.zeros(1, dtype=xd.dtype, device="cuda"); total = torch.zeros(1, dtype=xd.dtype, device="cuda")
            dr.ops.prefix_reduce_carry(OPS["add"], xd[:cut], True, False, carry_in=None, total_out=carry, vt=VT[vt], out=out[:cut])
            dr.ops.prefix_reduce_carry(OPS["add"], xd[cut:], True, False, carry_in=carry, total_out=total, vt=VT[vt], out=out[cut:])
            assert np.array_equal(to_np(out, vt), exp), (vt, n, cut)
            assert to_np(total, vt)[0] == (x.astype(np.uint64).sum() & np.uint64((1 << (8 * item)) - 1)).astype(x.dtype), (vt, n)


@pytest.mark.parametrize("vt", ["u8", "u32", "u64", "i32", "i64"])
def test_prefix_windowed_carry_many_tiles(vt):
    """Unsegmented scans on the TMA path with more tiles than CTAs (every CTA advances its carry
    by a full window of 296 aggregates several times), all four variants, ragged last tile; exact."""
    n = (1 << 23) + 12345 if vt != "u8" else (1 << 24) + 77
    x = make_input(vt, n); xd = to_dev(x, vt)
    for op in ("add", "max"):
        for ex, rev in ((1, 0), (0, 0), (1, 1)):
            if rev and n % (16 // x.itemsize):
                continue        # (mirrored vectors need the array end on a vector boundary: slow path, tested elsewhere)
            got = to_np(dr.block_prefix_reduce(OPS[op], xd, n, ex, rev, vt=VT[vt]), vt)
            assert np.array_equal(got, capi.block_prefix_reduce(vt, op, x, n, ex, rev)), (vt, op, ex, rev)
    n2 = n - n % 16
    got = to_np(dr.block_prefix_reduce(OPS["add"], xd[:n2], n2, 1, 1, vt=VT[vt]), vt)
    assert np.array_equal(got, capi.block_prefix_reduce(vt, "add", x[:n2], n2, 1, 1))


@pytest.mark.parametrize("vt", ["f32", "f64"])
def test_prefix_sum_float_large_and_reproducible(vt):
    """f32: relative 1e-6*log2(N) against an f64 oracle (north_star); the windowed carry fixes
    the association order, so two runs must agree bit for bit."""
    n = (1 << 24) + 3
    x = make_input(vt, n); xd = to_dev(x, vt)
    a = dr.block_prefix_reduce(ReduceOp.Add, xd, n, False, False, vt=VT[vt])
    b = dr.block_prefix_reduce(ReduceOp.Add, xd, n, False, False, vt=VT[vt])
    assert torch.equal(a, b)
    exp = np.cumsum(x.astype(np.float64))
    assert_close(to_np(a, vt), exp, vt, n, f"{vt} inclusive prefix sum of {n}")


# --------------------------------------------------------------------------- compress
def test_compress_reference_grid():
    """reductions.cpp:269-313: sizes 23 i^3 + 1, n_ones 23 j^3 + 1 random ones, exact list + count"""
    rng = np.random.default_rng(0)
    for i in range(0, 30, 2):
        size = 23 * i ** 3 + 1
        for j in range(0, i + 1, 3):
            data = np.zeros(size, np.uint8)
            data[rng.integers(0, size, 23 * j ** 3 + 1)] = 1
            got = to_np(dr.compress(to_dev(data, "u8")), "u32")
            assert np.array_equal(got, np.flatnonzero(data).astype(np.uint32)), (size, j)


@pytest.mark.parametrize("size", [1, 4095, 4096, 4097, 8192, 100000, 200001])
@pytest.mark.parametrize("density", [0.0, 0.01, 0.5, 0.99, 1.0])
def test_compress_large(size, density):
    """tests/test_reduction.py:401-415 (same seeds), bool masks, misaligned variant too"""
    rng = np.random.default_rng(seed=0xC0FFEE ^ size)
    mask = rng.uniform(0.0, 1.0, size) < density
    exp = np.flatnonzero(mask).astype(np.uint32)
    m = torch.from_numpy(mask).cuda()
    assert np.array_equal(to_np(dr.compress(m), "u32"), exp)
    buf = torch.zeros(size + 5, dtype=torch.bool, device="cuda")
    buf[5:] = m
    assert np.array_equal(to_np(dr.compress(buf[5:]), "u32"), exp)
    # the mask must not be modified (the reference zero-pads it, we do not)
    assert torch.equal(buf[5:], m)


@pytest.mark.parametrize("density", [0.003, 0.3, 0.97])
def test_compress_many_tiles_ragged_unaligned(density):
    """More tiles than CTAs (windowed carry over full windows), ragged tail, TMA and direct
    (unaligned mask) paths, sparse- and dense-row expansion; exact."""
    n = (1 << 25) + 4321
    rng = np.random.default_rng(7)
    m = (rng.random(n + 3) < density).astype(np.uint8)
    m[::977] *= 7                    # any non-zero byte counts (cuda_ts.cpp:683-763 treats the mask as bool)
    md = torch.from_numpy(m).cuda()
    for off in (0, 3):
        got = to_np(dr.compress(md[off:off + n]), "u32")
        assert np.array_equal(got, np.flatnonzero(m[off:off + n]).astype(np.uint32)), (density, off)


def test_compress_literals_and_golden():
    a = torch.tensor([0, 1, 1, 0, 0, 1, 0, 1, 1], dtype=torch.bool, device="cuda")
    assert dr.compress(a).tolist() == [1, 2, 5, 7, 8]
    assert dr.compress(torch.zeros(3, dtype=torch.bool, device="cuda")).tolist() == []
    assert dr.compress(torch.zeros(0, dtype=torch.bool, device="cuda")).tolist() == []
    digests, _ = loader.load()
    for n in [1, 9, 4095, 4096, 4097, 8192, 20001]:
        for thr in [0, 3, 128, 253, 256]:
            got = to_np(dr.compress(to_dev(capi.mask_u8(n, thr), "u8")), "u32")
            assert digest(got)[0] == digests[f"compress/{n}/{thr}"]


def test_compress_2_30_properties():
    """BASELINE size 2^30, density 50 %: count == popcount, strictly increasing, mask[out] all set,
    first 2^22 indices bit-exact against the oracle"""
    n = 1 << 30
    m = torch.empty(n, dtype=torch.uint8, device="cuda")
    dr.ops.fill_fmix32(m, 2, and_=128)
    out = dr.compress(m)
    count = out.numel()
    assert count == int(m.sum(dtype=torch.int64).item())
    o = out.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    assert bool(torch.all(o[1:] > o[:-1]))
    assert bool(torch.all(m[o] == 1))
    exp = capi.compress(capi.mask_u8(1 << 22, 128))
    assert np.array_equal(o[:exp.size].cpu().numpy().astype(np.uint32), exp)


# --------------------------------------------------------------------------- mkperm
def check_mkperm(keys, buckets, perm, table, block_size=None):
    """The reference's own acceptance test (reductions.cpp:360-398): ids / starts / sizes of
    the non-empty buckets and the *sorted* contents of every bucket."""
    n = keys.size
    bs = block_size or n
    exp_perm, exp_off, exp_unique = capi.block_mkperm(keys, bs, buckets)
    if table is not None:
        t = table.numpy().astype(np.uint32)
        assert t.shape[0] == exp_unique
        assert np.array_equal(t.reshape(-1), exp_off[:4 * exp_unique])
    # per bucket (and per group) the contents must agree as sets; sorting each run of equal
    # keys makes an unstable-but-valid permutation comparable with the stable oracle
    got = perm.astype(np.int64)
    assert np.array_equal(np.sort(got), np.arange(n)), "not a permutation"
    group = np.arange(n) // bs
    key_of = keys[got].astype(np.int64)
    assert np.array_equal(group[got], group), "element left its sorting group"
    order_key = group * (buckets + 1) + key_of
    assert np.all(np.diff(order_key) >= 0), "keys not sorted inside groups"
    canon = np.lexsort((got, order_key))
    assert np.array_equal(got[canon], exp_perm.astype(np.int64))


def test_mkperm_reference_grid():
    """reductions.cpp:315-406 (sizes / bucket counts 23 i^3 + 1)"""
    rng = np.random.default_rng(0)
    for i in range(0, 30, 3):
        size = 23 * i ** 3 + 1
        for j in range(0, i + 1, 3):
            buckets = 23 * j ** 3 + 1
            keys = rng.integers(0, buckets, size).astype(np.uint32)
            perm, table = dr.block_mkperm(to_dev(keys, "u32"), size, buckets)
            torch.cuda.synchronize()
            check_mkperm(keys, buckets, to_np(perm, "u32"), table)


def test_mkperm_stable_variant_is_bit_exact():
    """bucket counts that fit the per-warp variant must reproduce the reference's CPU
    (stable) permutation bit for bit, including the golden digests"""
    digests, _ = loader.load()
    for n, buckets in [(1, 1), (24, 1), (185, 24), (622, 185), (1473, 622), (20001, 4096), (20001, 37)]:
        keys = capi.fmix32(n) % np.uint32(buckets)
        perm, table = dr.block_mkperm(to_dev(keys, "u32"), n, buckets)
        torch.cuda.synchronize()
        assert digest(to_np(perm, "u32"))[0] == digests[f"mkperm/{n}/{buckets}/perm"], (n, buckets)
        assert digest(table.numpy().astype(np.uint32).reshape(-1))[0] == digests[f"mkperm/{n}/{buckets}/offsets"]
        for bs in (7, 256):
            if bs < n:
                perm, table = dr.block_mkperm(to_dev(keys, "u32"), bs, buckets)
                torch.cuda.synchronize()
                assert table is None
                assert digest(to_np(perm, "u32"))[0] == digests[f"mkperm_block/{n}/{buckets}/{bs}"], (n, buckets, bs)


@pytest.mark.parametrize("buckets", [1, 2, 37, 256, 512, 513, 1000, 1816])
@pytest.mark.parametrize("n", [(1 << 18) + 5, (1 << 21) + 12345])
def test_mkperm_large_inputs_stay_stable_where_the_reference_is(buckets, n):
    """jit.h:2404-2406: while bucket_count * 4 B * 32 warps fit into shared memory (<= 1816 buckets
    here) the reference's permutation is stable; the tile path used for large inputs must
    reproduce the stable (CPU reference) permutation bit for bit, table included."""
    keys = capi.fmix32(n) % np.uint32(buckets)
    if buckets > 2:
        keys[keys == 1] = 0                      # one empty bucket
    perm, table = dr.block_mkperm(to_dev(keys, "u32"), n, buckets)
    torch.cuda.synchronize()
    eperm, eoff, eunique = capi.block_mkperm(keys, n, buckets)
    assert np.array_equal(to_np(perm, "u32"), eperm)
    assert table.shape[0] == eunique and np.array_equal(table.numpy().astype(np.uint32).reshape(-1), eoff[:4 * eunique])


def test_lsd_radix_sort_on_top_of_block_mkperm():
    """dr.sort / dr.argsort (drjit/__init__.py:1698-1772): LSD radix sort of 32-bit keys made of
    four block_mkperm passes with 256 buckets; only correct if every pass is stable."""
    n = (1 << 20) + 777
    x = torch.from_numpy(capi.fmix32(n, xor=0x1234567).view(np.int32)).cuda()
    ordinal = x.to(torch.int64) & 0xFFFFFFFF
    ordinal[::3] &= 0xFFFF0000                   # plenty of equal keys: the result must be the *stable* order
    index = torch.arange(n, device="cuda")
    cur, cur_index = ordinal.clone(), index.clone()
    for shift in (0, 8, 16, 24):
        digit = ((cur >> shift) & 0xFF).to(torch.int32)
        perm, _ = dr.block_mkperm(digit, n, 256, want_offsets=False)
        torch.cuda.synchronize()
        pl = perm.to(torch.int64)
        cur, cur_index = cur[pl], cur_index[pl]
    exp_sorted, exp_index = torch.sort(ordinal, stable=True)
    assert torch.equal(cur, exp_sorted) and torch.equal(cur_index, exp_index)


@pytest.mark.parametrize("buckets", [1, 2, 37, 256, 4096, 8000, 50000, 100000])
def test_mkperm_variants(buckets):
    """per-warp (stable), per-CTA and global-atomic variants; single group and sorting groups"""
    n = 1_000_003
    keys = capi.fmix32(n) % np.uint32(buckets)
    kd = to_dev(keys, "u32")
    perm, table = dr.block_mkperm(kd, n, buckets)
    torch.cuda.synchronize()
    check_mkperm(keys, buckets, to_np(perm, "u32"), table)
    for bs in (1000, 65536):
        perm, table = dr.block_mkperm(kd, bs, buckets)
        torch.cuda.synchronize()
        check_mkperm(keys, buckets, to_np(perm, "u32"), None, block_size=bs)
    kd1 = to_dev(keys, "u32", 1)   # misaligned keys
    perm, table = dr.block_mkperm(kd1, n, buckets)
    torch.cuda.synchronize()
    check_mkperm(keys, buckets, to_np(perm, "u32"), table)


def check_mkperm_device(kd, buckets, perm, table):
    """Device-side form of the reference's acceptance test (reductions.cpp:360-398) for sizes
    where numpy would take minutes: `perm` is a permutation, keys[perm] is non-decreasing (so
    bucket b holds exactly the indices with key b), and the table of non-empty buckets matches
    a bincount of the keys."""
    n = kd.numel()
    p64 = perm.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    assert torch.equal(torch.sort(p64).values, torch.arange(n, device=kd.device)), "not a permutation"
    k64 = kd.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    gathered = k64[p64]
    assert bool(torch.all(gathered[1:] >= gathered[:-1])), "keys not sorted"
    counts = torch.bincount(k64, minlength=buckets)
    starts = torch.cumsum(counts, 0) - counts
    ids = torch.nonzero(counts).flatten()
    exp = torch.stack([ids, starts[ids], counts[ids], torch.zeros_like(ids)], dim=1).cpu()
    assert table.shape == exp.shape and torch.equal(table, exp)


def test_mkperm_baseline_config():
    """BASELINE config: 2^26 keys, 4096 buckets; uniform and skewed ids; plus one bucket only"""
    n = 1 << 26
    kd = torch.empty(n, dtype=torch.int32, device="cuda")
    for variant in ("uniform", "skewed", "single"):
        dr.ops.fill_fmix32(kd, 0, and_=4095)
        if variant == "skewed":
            other = torch.empty_like(kd)
            dr.ops.fill_fmix32(other, 0, xor=0x9E3779B9, and_=4095)
            kd = torch.minimum(kd, other)
        elif variant == "single":
            kd.fill_(1234)
        perm, table = dr.block_mkperm(kd, n, 4096)
        torch.cuda.synchronize()
        check_mkperm_device(kd, 4096, perm, table)
    # the first 2^20 keys against the oracle (sets per bucket)
    keys = capi.fmix32(1 << 20, mask=4095)
    perm, table = dr.block_mkperm(to_dev(keys, "u32"), 1 << 20, 4096)
    torch.cuda.synchronize()
    check_mkperm(keys, 4096, to_np(perm, "u32"), table)


@pytest.mark.parametrize("buckets", [1817, 4096, 4352, 4353, 8192])
def test_mkperm_unordered_tiles_ragged_unaligned(buckets):
    """unordered tile kernel (beyond the reference's stable range) at sizes where the largest tile
    applies: ragged last tile, input not 16-byte aligned, keys drawn from all / from few buckets"""
    import os
    kpt = max(48, int(os.environ.get("DRJIT_B200_MKPERM_KPT", "48") or 48))     # (A/B runs of larger tiles)
    n = 148 * 2 * kpt * 1024 + 12_345
    buf = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    for mis in (0, 1):
        kd = buf[mis:mis + n]
        dr.ops.fill_fmix32(kd, 0, xor=buckets)
        kd.remainder_(buckets)
        perm, table = dr.block_mkperm(kd, n, buckets)
        torch.cuda.synchronize()
        check_mkperm_device(kd, buckets, perm, table)
    kd = buf[:n]
    kd.remainder_(3).mul_(buckets // 3)                   # three buckets hold everything
    perm, table = dr.block_mkperm(kd, n, buckets)
    torch.cuda.synchronize()
    check_mkperm_device(kd, buckets, perm, table)


def test_mkperm_errors():
    k = torch.zeros(10, dtype=torch.int32, device="cuda")
    with pytest.raises(dr._lib.FatalError, match="bucket_count cannot be zero"):
        dr.block_mkperm(k, 10, 0)
    perm, table = dr.block_mkperm(k[:0], 0, 5)
    assert perm.numel() == 0


# --------------------------------------------------------------------------- scatter_reduce
@pytest.mark.parametrize("mode", [ReduceMode.Auto, ReduceMode.Direct, ReduceMode.Local])
def test_scatter_reduce_modes_agree(mode):
    """tests/test_memop.py:423-499: all modes must agree, target sizes 2^0..2^9 (+ larger)"""
    n = 100_003
    for k in list(range(0, 10)) + [16, 20]:
        bins = 1 << k
        idx = capi.fmix32(n, xor=0x85EBCA6B, mask=bins - 1)
        for vt in ("u32", "f32", "f64", "i64"):
            val = make_input(vt, n)
            if vt in ("u32", "i64"):
                val = (val.astype(np.int64) % 1000).astype(NP[vt])
            tgt = np.zeros(bins, NP[vt])
            exp = capi.scatter_reduce(vt, "add", tgt, val, idx, acc64=True)
            got = to_np(dr.scatter_reduce(ReduceOp.Add, to_dev(tgt, vt), to_dev(val, vt), to_dev(idx, "u32"),
                                          mode=mode, vt=VT[vt]), vt)
            if vt in ("u32", "i64"):
                assert np.array_equal(got, exp), (k, vt)
            else:
                assert_close(got, exp, vt, max(2, n // bins), f"scatter add {vt} bins=2^{k}")


@pytest.mark.parametrize("op", ["min", "max", "and", "or"])
def test_scatter_reduce_other_ops(op):
    n, bins = 50_001, 97
    idx = capi.fmix32(n, xor=1) % np.uint32(bins)
    mask = (capi.fmix32(n, xor=2) & 3) != 0
    for vt in ("u32", "i32", "u64", "i64", "f32", "f64"):
        if op in ("and", "or") and vt in ("f32", "f64"):
            with pytest.raises(RuntimeError, match="does not support"):
                dr.scatter_reduce(OPS[op], torch.zeros(4, device="cuda"), torch.zeros(4, device="cuda"),
                                  torch.zeros(4, dtype=torch.int32, device="cuda"))
            continue
        val = make_input(vt, n)
        init = make_input(vt, bins)
        exp = capi.scatter_reduce(vt, op, init, val, idx, mask=mask.astype(np.uint8))
        for mode in (ReduceMode.Direct, ReduceMode.Local):
            got = to_np(dr.scatter_reduce(OPS[op], to_dev(init, vt), to_dev(val, vt), to_dev(idx, "u32"),
                                          active=torch.from_numpy(mask).cuda(), mode=mode, vt=VT[vt]), vt)
            assert np.array_equal(got, exp), (vt, op, mode)


@pytest.mark.parametrize("vt", ["f32", "f64"])
def test_scatter_minmax_negative_zero(vt):
    """-0.0 has the sign bit set, so the reference sends it down the unsigned-atomic path
    (`setp.ge.s32` on the bit pattern, src/cuda_scatter.cpp:98-105): min(-5, -0.0) = -5,
    max(-5, -0.0) = -0.0, min(+0.0, -0.0) = -0.0, max(-0.0, +0.0) = +0.0 -- the IEEE total order,
    compared bit for bit."""
    npt = NP[vt]
    bits = np.uint32 if vt == "f32" else np.uint64
    init = np.array([-5.0, -5.0, 0.0, -0.0, 3.0, -1e-30, -0.0, 7.0], npt)
    val = np.array([-0.0] * 6 + [0.0, -0.0], npt)
    idx = np.arange(8, dtype=np.uint32)

    def total_order_key(a):     # monotone map of IEEE bit patterns to unsigned integers
        b = a.view(bits).astype(np.uint64)
        sign = np.uint64(1) << np.uint64(8 * a.itemsize - 1)
        allones = np.uint64((1 << (8 * a.itemsize)) - 1)
        return np.where(b & sign, (~b) & allones, b | sign)

    for op in ("min", "max"):
        ka, kb = total_order_key(init), total_order_key(val)
        take_val = (kb < ka) if op == "min" else (kb > ka)
        exp = np.where(take_val, val, init).astype(npt)
        for mode in (ReduceMode.Direct, ReduceMode.Local):
            got = to_np(dr.scatter_reduce(OPS[op], to_dev(init, vt), to_dev(val, vt), to_dev(idx, "u32"),
                                          mode=mode, vt=VT[vt]), vt)
            assert np.array_equal(got.view(bits), exp.view(bits)), (vt, op, mode, got, exp)


@pytest.mark.parametrize("op", ["add", "min", "max"])
@pytest.mark.parametrize("misalign", [0, 1])
def test_scatter_reduce_f16(op, misalign):
    """f16 scatter-reductions: add since cc 60, min/max since cc 90 (src/op.cpp:2781-2794); the CUDA
    form is the two-wide f16 reduction with an identity partner (src/cuda_scatter.cpp:291-332).
    Sums of small integers are exact in f16, so every comparison here is bit-for-bit (-0 == +0)."""
    n, bins = 40_003, 501           # odd bin count: the last element has no partner inside the array
    idx = capi.fmix32(n, xor=3) % np.uint32(bins)
    mask = (capi.fmix32(n, xor=4) & 7) != 0
    if op == "add":
        val = (capi.fmix32(n, xor=5) % np.uint32(8)).astype(np.float16)     # bin sums < 2048: exact
        init = (capi.fmix32(bins, xor=6) % np.uint32(16)).astype(np.float16)
    else:
        val = make_input("f16", n)
        init = make_input("f16", bins)
    exp = capi.scatter_reduce("f16", op, init, val, idx, mask=mask.astype(np.uint8))
    for mode in (ReduceMode.Auto, ReduceMode.Direct, ReduceMode.Local):
        tgt = to_dev(init, "f16", misalign=misalign)
        got = to_np(dr.scatter_reduce(OPS[op], tgt, to_dev(val, "f16"), to_dev(idx, "u32"),
                                      active=torch.from_numpy(mask).cuda(), mode=mode), "f16")
        assert np.array_equal(got, exp), (op, mode, misalign)
    for bad in ("and", "or", "mul"):
        with pytest.raises(RuntimeError, match="does not support"):
            dr.scatter_reduce(OPS[bad], torch.zeros(4, dtype=torch.float16, device="cuda"),
                              torch.zeros(4, dtype=torch.float16, device="cuda"),
                              torch.zeros(4, dtype=torch.int32, device="cuda"))


def test_scatter_add_baseline_config():
    """BASELINE config 5 (one shard): 2^25 f32 values into 2^20 bins; tolerance 1e-6*log2(count)"""
    n, bins = 1 << 25, 1 << 20
    val = capi.unit_f32(n)
    idx = capi.fmix32(n, xor=0x85EBCA6B, mask=bins - 1)
    exp = capi.scatter_reduce("f32", "add", np.zeros(bins, np.float32), val, idx, acc64=True)
    got = to_np(dr.scatter_add(torch.zeros(bins, device="cuda"), to_dev(val, "f32"), to_dev(idx, "u32")), "f32")
    assert_close(got, exp, "f32", 32, "histogram")


# --------------------------------------------------------------------------- misc
def test_memset_poke_aggregate():
    from drjit_b200._lib import check, lib
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for isize, pattern in [(1, b"\x07"), (2, b"\x34\x12"), (4, b"\x78\x56\x34\x12"), (8, bytes(range(1, 9))),
                           (4, b"\0\0\0\0"), (8, b"\xff" * 8)]:
        for n in (1, 3, 1000, 100_001):
            for mis in (0, 1):
                buf = torch.zeros((n + mis) * isize + 16, dtype=torch.uint8, device="cuda")
                view = buf[mis * isize:(mis + n) * isize]
                dr.ops.memset(view, pattern)
                exp = np.frombuffer(pattern * n, np.uint8)
                assert np.array_equal(view.cpu().numpy(), exp), (isize, n, mis)
                assert int(buf[(mis + n) * isize:].sum().item()) == 0 and int(buf[:mis * isize].sum().item()) == 0
    with pytest.raises(RuntimeError, match="invalid element size"):
        dr.ops.memset(torch.zeros(12, dtype=torch.uint8, device="cuda"), b"\1\2\3")
    # poke
    buf = torch.zeros(4, dtype=torch.int64, device="cuda")
    v = ctypes.c_uint64(0x1122334455667788)
    check(lib.drjit_b200_poke(stream, ctypes.c_void_p(buf.data_ptr() + 8), ctypes.byref(v), 8))
    v4 = ctypes.c_uint32(0xDEADBEEF)
    check(lib.drjit_b200_poke(stream, ctypes.c_void_p(buf.data_ptr() + 16), ctypes.byref(v4), 4))
    assert buf.tolist() == [0, 0x1122334455667788, 0xDEADBEEF, 0]
    # aggregate (resources/misc.cuh:41-61): literals (size > 0) and dereferenced sources (size < 0)
    src = torch.tensor([0x0102030405060708], dtype=torch.int64, device="cuda")
    entries = np.zeros(4, dtype=[("size", "<i2"), ("kind", "<u2"), ("offset", "<u4"), ("src", "<u8")])
    entries[0] = (4, 0, 0, 0xAABBCCDD)
    entries[1] = (-8, 0, 8, src.data_ptr())
    entries[2] = (1, 0, 4, 0x7F)
    entries[3] = (-2, 0, 6, src.data_ptr())
    agg = torch.from_numpy(entries.view(np.uint8).copy()).cuda()
    dst = torch.zeros(16, dtype=torch.uint8, device="cuda")
    check(lib.drjit_b200_aggregate(stream, ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(agg.data_ptr()), 4))
    exp = np.zeros(16, np.uint8)
    exp[0:4] = np.frombuffer((0xAABBCCDD).to_bytes(4, "little"), np.uint8)
    exp[4] = 0x7F
    exp[6:8] = [0x08, 0x07]
    exp[8:16] = np.frombuffer((0x0102030405060708).to_bytes(8, "little"), np.uint8)
    assert np.array_equal(dst.cpu().numpy(), exp)


def test_launch_accounting_and_native_library_loaded():
    """the CUDA path is the one that runs: launches are counted and the .so is mapped"""
    dr.launch_count(reset=True)
    x = torch.ones(1 << 20, device="cuda")
    dr.sum(x); dr.prefix_sum(x)
    assert dr.launch_count() >= 2
    with open("/proc/self/maps") as f:
        assert "libdrjit_b200.so" in f.read()
