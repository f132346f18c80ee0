"""Property tests of the CPU oracle (oracle/oracle.c) against independent numpy restatements of the
reference semantics, over random shapes -- hypothesis picks (size, block size, op, type, flags). The
known-answer / golden / live-reference pins are in tests/test_oracle.py; this file guards the oracle's
handling of ragged blocks, wrap-around and the exclusive / reverse variants at sizes no fixture lists.
Semantics restated from: llvm_red.h:87-179 (block reductions and prefix reductions, identities
var.cpp:2642-2652), llvm_ts.cpp:706-780 (compress), llvm_ts.cpp:785-933 (stable mkperm + offsets table),
jit.h:1076-1105 (scatter-reduce)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import capi

INT_TYPES = {"u32": np.uint32, "i32": np.int32, "u64": np.uint64, "i64": np.int64}
OPS = ["add", "mul", "min", "max", "and", "or"]
COMMON = dict(max_examples=60, deadline=None)


def np_reduce(op, a, dtype):
    """Reduction of a 1-D integer array with wrap-around in `dtype`"""
    if a.size == 0:
        raise ValueError
    with np.errstate(over="ignore"):
        return {"add": lambda: np.add.reduce(a, dtype=dtype), "mul": lambda: np.multiply.reduce(a, dtype=dtype),
                "min": lambda: a.min(), "max": lambda: a.max(),
                "and": lambda: np.bitwise_and.reduce(a), "or": lambda: np.bitwise_or.reduce(a)}[op]()


def np_accumulate(op, a, dtype):
    with np.errstate(over="ignore"):
        return {"add": lambda: np.add.accumulate(a, dtype=dtype), "mul": lambda: np.multiply.accumulate(a, dtype=dtype),
                "min": lambda: np.minimum.accumulate(a), "max": lambda: np.maximum.accumulate(a),
                "and": lambda: np.bitwise_and.accumulate(a), "or": lambda: np.bitwise_or.accumulate(a)}[op]()


def identity(vt, op):
    dt = INT_TYPES[vt]
    info = np.iinfo(dt)
    return {"add": dt(0), "mul": dt(1), "min": dt(info.max), "max": dt(info.min),
            "and": dt(info.max) if info.min == 0 else dt(-1), "or": dt(0)}[op]


def random_array(seed, n, vt):
    raw = capi.fmix32_u64(n, start=seed * 7919)
    return (raw * np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64).view(np.uint64).astype(INT_TYPES[vt], casting="unsafe") \
        if vt in ("u64", "i64") else raw.astype(np.uint32).view(INT_TYPES[vt])


@settings(**COMMON)
@given(seed=st.integers(0, 1000), n=st.integers(1, 700), bs=st.integers(1, 700), vt=st.sampled_from(sorted(INT_TYPES)),
       op=st.sampled_from(OPS))
def test_block_reduce_matches_numpy(seed, n, bs, vt, op):
    bs = min(bs, n)
    x = random_array(seed, n, vt)
    got = capi.block_reduce(vt, op, x, bs)
    exp = np.array([np_reduce(op, x[b:b + bs], INT_TYPES[vt]) for b in range(0, n, bs)], INT_TYPES[vt])
    assert np.array_equal(got, exp)


@settings(**COMMON)
@given(seed=st.integers(0, 1000), n=st.integers(1, 500), bs=st.integers(1, 500), vt=st.sampled_from(sorted(INT_TYPES)),
       op=st.sampled_from(OPS), exclusive=st.booleans(), reverse=st.booleans())
def test_block_prefix_reduce_matches_numpy(seed, n, bs, vt, op, exclusive, reverse):
    bs = min(bs, n)
    dt = INT_TYPES[vt]
    x = random_array(seed, n, vt)
    exp = np.empty_like(x)
    for b in range(0, n, bs):
        blk = x[b:b + bs]
        if reverse:
            blk = blk[::-1]
        inc = np_accumulate(op, blk, dt).astype(dt)
        res = np.concatenate(([identity(vt, op)], inc[:-1])).astype(dt) if exclusive else inc
        exp[b:b + bs] = res[::-1] if reverse else res
    got = capi.block_prefix_reduce(vt, op, x, bs, exclusive, reverse)
    assert np.array_equal(got, exp)


@settings(**COMMON)
@given(seed=st.integers(0, 1000), n=st.integers(0, 5000), threshold=st.integers(0, 256))
def test_compress_matches_nonzero(seed, n, threshold):
    mask = capi.mask_u8(n, threshold, start=seed)
    assert np.array_equal(capi.compress(mask), np.flatnonzero(mask).astype(np.uint32))


@settings(**COMMON)
@given(seed=st.integers(0, 1000), n=st.integers(1, 3000), buckets=st.integers(1, 300))
def test_mkperm_is_the_stable_sort_and_lists_the_non_empty_buckets(seed, n, buckets):
    keys = capi.fmix32(n, start=seed * 31) % np.uint32(buckets)
    perm, offsets, unique = capi.block_mkperm(keys, n, buckets)
    assert np.array_equal(perm, np.argsort(keys, kind="stable").astype(np.uint32))
    ids, counts = np.unique(keys, return_counts=True)
    assert unique == ids.size
    table = offsets[:4 * unique].reshape(-1, 4)
    starts = np.concatenate(([0], np.cumsum(counts)[:-1]))
    assert np.array_equal(table[:, 0], ids) and np.array_equal(table[:, 1], starts) and np.array_equal(table[:, 2], counts)
    assert offsets[4 * buckets] == unique                           # unique count behind the table (llvm_ts.cpp:918-924)


@settings(**COMMON)
@given(seed=st.integers(0, 1000), n=st.integers(1, 2000), bs_pow=st.integers(0, 6), buckets=st.integers(1, 40))
def test_mkperm_blocks_are_sorted_independently(seed, n, bs_pow, buckets):
    bs = min(n, 1 << bs_pow << 3)
    keys = capi.fmix32(n, start=seed * 17) % np.uint32(buckets)
    perm, _, _ = capi.block_mkperm(keys, bs, buckets, want_offsets=False)
    for b in range(0, n, bs):
        blk = keys[b:b + bs]
        assert np.array_equal(perm[b:b + bs], b + np.argsort(blk, kind="stable").astype(np.uint32))


@settings(**COMMON)
@given(seed=st.integers(0, 1000), n=st.integers(0, 3000), bins=st.integers(1, 200), vt=st.sampled_from(sorted(INT_TYPES)),
       op=st.sampled_from(["add", "min", "max", "and", "or"]), masked=st.booleans())
def test_scatter_reduce_matches_ufunc_at(seed, n, bins, vt, op, masked):
    dt = INT_TYPES[vt]
    val = random_array(seed, n, vt)
    idx = capi.fmix32(n, start=seed * 13 + 1) % np.uint32(bins)
    mask = (capi.fmix32(n, start=seed * 13 + 2) & 1).astype(np.uint8) if masked else None
    init = random_array(seed + 1, bins, vt)
    exp = init.copy()
    sel = mask.astype(bool) if masked else np.ones(n, bool)
    ufunc = {"add": np.add, "min": np.minimum, "max": np.maximum, "and": np.bitwise_and, "or": np.bitwise_or}[op]
    with np.errstate(over="ignore"):
        ufunc.at(exp, idx[sel], val[sel])
    got = capi.scatter_reduce(vt, op, init, val, idx, mask=mask)
    assert got.dtype == dt and np.array_equal(got, exp)


@settings(**COMMON)
@given(seed=st.integers(0, 1000), n=st.integers(0, 3000), counters=st.integers(1, 64), masked=st.booleans())
def test_scatter_inc_hands_out_consecutive_slots_in_element_order(seed, n, counters, masked):
    idx = capi.fmix32(n, start=seed * 5) % np.uint32(counters)
    mask = (capi.fmix32(n, start=seed * 5 + 1) & 1).astype(np.uint8) if masked else None
    before = capi.fmix32(counters, start=seed) % np.uint32(100)
    after, out = capi.scatter_inc(before, idx, mask)
    sel = mask.astype(bool) if masked else np.ones(n, bool)
    assert np.array_equal(after, before + np.bincount(idx[sel], minlength=counters).astype(np.uint32))
    running = before.copy()
    for i in range(n):
        if sel[i]:
            assert out[i] == running[idx[i]]
            running[idx[i]] += 1
        else:
            assert out[i] == 0
