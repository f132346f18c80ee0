"""GPU parity tests of the packet scatter-reduce and of scatter_inc (scatter_packet.cu) against the CPU
oracle, the reference's own test properties (tests/test_memop.py:293-316 for scatter_inc) and -- when
oracle/_ref is present -- the unmodified reference's CUDA JIT on the same device buffers
(jit_var_scatter_packet / jit_var_scatter_inc through oracle/ref_build/ref_shim.cpp).

Integer results: bit-exact. f32 / f64 / f16 Add: the order of the atomics is unspecified on both sides,
tolerance 1e-6 * log2(entries per bin) relative for f32 (north_star), 1e-14-scale f64, 1.5e-3-scale f16.
Min / Max are order-independent: exact."""
import ctypes

import numpy as np
import pytest
import torch

import drjit_b200 as dr
from drjit_b200 import ReduceMode, ReduceOp
from oracle import capi, ref
from tests.gpu_util import NP, OPS, VT, to_dev, to_np
from tests.golden.make_golden import make_input
from tests.test_gpu_parity import assert_close

pytestmark = pytest.mark.gpu
vp = ctypes.c_void_p


def packet_inputs(vt, n, count, bins, small_ints=True):
    idx = capi.fmix32(n, xor=0x85EBCA6B) % np.uint32(bins)
    vals = []
    for k in range(count):
        v = make_input(vt, n + k)[k:]
        if small_ints and vt in ("u32", "i32", "u64", "i64"):
            v = (v.astype(np.int64) % 1000).astype(NP[vt])
        if vt == "f16":     # small integers: f16 sums stay exact whatever the order (as tests/test_gpu_parity.py does)
            v = (capi.fmix32(n, xor=20 + k) % np.uint32(8)).astype(np.float16) - np.float16(3 * (k & 1))
        vals.append(np.ascontiguousarray(v))
    return idx, vals


# --------------------------------------------------------------------------- packet scatter-reduce
@pytest.mark.parametrize("count", [2, 4, 6, 8, 16])
@pytest.mark.parametrize("mode", [ReduceMode.Auto, ReduceMode.Direct, ReduceMode.Local])
def test_packet_add_modes_agree(count, mode):
    """tests/test_memop.py:423-499 (all modes agree) on packets; bins from one packet to 2^16"""
    n = 60_003
    for bins in (1, 3, 64, 1 << 10, 1 << 16):
        for vt in ("u32", "f32", "f64", "i64"):
            idx, vals = packet_inputs(vt, n, count, bins)
            tgt = np.zeros(bins * count, NP[vt])
            exp = capi.scatter_reduce_packet(vt, "add", tgt, vals, idx, acc64=True)
            got = to_np(dr.scatter_reduce(ReduceOp.Add, to_dev(tgt, vt), [to_dev(v, vt) for v in vals],
                                          to_dev(idx, "u32"), mode=mode, vt=VT[vt]), vt)
            if vt in ("u32", "i64"):
                assert np.array_equal(got, exp), (bins, vt)
            else:
                assert_close(got, exp, vt, max(2, n // bins), f"packet add {vt} x{count} bins={bins}")


@pytest.mark.parametrize("op", ["min", "max", "and", "or"])
def test_packet_other_ops_masked(op):
    n, bins, count = 40_001, 97, 4
    mask = (capi.fmix32(n, xor=2) & 3) != 0
    for vt in ("u32", "i32", "u64", "i64", "f32", "f64"):
        idx, vals = packet_inputs(vt, n, count, bins, small_ints=False)
        if op in ("and", "or") and vt in ("f32", "f64"):
            with pytest.raises(RuntimeError, match="does not support"):
                dr.scatter_reduce(OPS[op], to_dev(np.zeros(bins * count, NP[vt]), vt),
                                  [to_dev(v, vt) for v in vals], to_dev(idx, "u32"))
            continue
        init = make_input(vt, bins * count)
        exp = capi.scatter_reduce_packet(vt, op, init, vals, idx, mask=mask.astype(np.uint8))
        for mode in (ReduceMode.Direct, ReduceMode.Local):
            got = to_np(dr.scatter_reduce(OPS[op], to_dev(init, vt), [to_dev(v, vt) for v in vals],
                                          to_dev(idx, "u32"), active=torch.from_numpy(mask).cuda(), mode=mode,
                                          vt=VT[vt]), vt)
            assert np.array_equal(got, exp), (vt, op, mode)


@pytest.mark.parametrize("op", ["add", "min", "max"])
@pytest.mark.parametrize("count", [2, 4, 8, 12])
def test_packet_f16(op, count):
    """f16 packets leave as red.global.v{2,4,8}.f16 (cuda_packet.cpp:224-259)"""
    n, bins = 20_000, 4099
    idx, vals = packet_inputs("f16", n, count, bins)
    mask = (capi.fmix32(n, xor=9) & 7) != 0
    init = (capi.fmix32(bins * count, xor=6) % np.uint32(16)).astype(np.float16)
    exp = capi.scatter_reduce_packet("f16", op, init, vals, idx, mask=mask.astype(np.uint8))
    for misalign in (0, 2, 4):      # 16- / 4- / 8-byte aligned targets: v8 (v4, v2) / v2 / v4 reductions
        got = to_np(dr.scatter_reduce(OPS[op], to_dev(init, "f16", misalign), [to_dev(v, "f16") for v in vals],
                                      to_dev(idx, "u32"), active=torch.from_numpy(mask).cuda()), "f16")
        assert np.array_equal(got, exp), (op, count, misalign)      # (-0 == +0)


@pytest.mark.parametrize("misalign", [1, 2, 3])
def test_packet_unaligned_target_and_inputs(misalign):
    """A target that is not aligned to the vector falls back to narrower reductions; component /
    index arrays at odd offsets are read with scalar loads anyway."""
    n, bins, count = 10_007, 513, 4
    idx, vals = packet_inputs("f32", n, count, bins)
    tgt = np.zeros(bins * count, np.float32)
    exp = capi.scatter_reduce_packet("f32", "add", tgt, vals, idx, acc64=True)
    got = to_np(dr.scatter_reduce(ReduceOp.Add, to_dev(tgt, "f32", misalign),
                                  [to_dev(v, "f32", misalign) for v in vals], to_dev(idx, "u32", misalign)), "f32")
    assert_close(got, exp, "f32", max(2, n // bins), "unaligned packet add")


def test_packet_argument_errors():
    t = torch.zeros(12, device="cuda"); i = torch.zeros(4, dtype=torch.int32, device="cuda")
    v = [torch.zeros(4, device="cuda") for _ in range(3)]
    with pytest.raises(RuntimeError, match="not supported by reduction"):          # odd packet: cuda_packet.cpp:184-186
        dr.scatter_reduce(ReduceOp.Add, t, v, i)
    with pytest.raises(RuntimeError, match="not supported by reduction"):
        dr.scatter_reduce(ReduceOp.Add, torch.zeros(36, device="cuda"), [torch.zeros(4, device="cuda")] * 18, i)
    # empty input: no-op
    e = torch.zeros(0, device="cuda")
    out = dr.scatter_reduce(ReduceOp.Add, torch.ones(8, device="cuda"), [e, e], torch.zeros(0, dtype=torch.int32, device="cuda"))
    assert torch.equal(out, torch.ones(8, device="cuda"))


def test_packet_film_config():
    """2^22 RGBA samples into a 2^16-pixel film: every packet leaves as one REDG.E.ADD.F32x4"""
    n, bins, count = 1 << 22, 1 << 16, 4
    idx, vals = packet_inputs("f32", n, count, bins)
    exp = capi.scatter_reduce_packet("f32", "add", np.zeros(bins * count, np.float32), vals, idx, acc64=True)
    dr.launch_count(reset=True)
    got = to_np(dr.scatter_add(torch.zeros(bins * count, device="cuda"), [to_dev(v, "f32") for v in vals],
                               to_dev(idx, "u32")), "f32")
    assert dr.launch_count() == 1
    assert_close(got, exp, "f32", n // bins, "film accumulation")


# --------------------------------------------------------------------------- scatter_inc
def check_scatter_inc(counters_before, counters_after, index, mask, out, what):
    """The reference's own acceptance test (tests/test_memop.py:305-316) with initial counter values and
    a mask: counters advance by the histogram; the slots handed out for counter j are exactly
    before[j] .. before[j] + hist[j] - 1, each once; masked elements receive 0."""
    B = counters_before.size
    active = np.ones(index.size, bool) if mask is None else mask.astype(bool)
    hist = np.bincount(index[active], minlength=B).astype(np.uint32)
    assert np.array_equal(counters_after, counters_before + hist), what
    assert np.all(out[~active] == 0), what
    order = np.lexsort((out[active], index[active]))
    slots = out[active][order]; owner = index[active][order]
    starts = np.concatenate(([0], np.cumsum(hist)[:-1]))
    expected = counters_before[owner] + (np.arange(slots.size, dtype=np.uint32) - starts[owner].astype(np.uint32))
    assert np.array_equal(slots, expected), what


@pytest.mark.parametrize("B", list(range(1, 17)) + [100, 2048, 2049, 1 << 16])
def test_scatter_inc_reference_grid(B):
    """tests/test_memop.py:293-316: random increments into 2..16 counters, 10000 elements (+ larger
    counter arrays: the shared-memory path ends at 2048 counters)"""
    rng = np.random.RandomState(B)
    for n in (1, 31, 10_000, 100_003):
        index = rng.randint(0, B, n).astype(np.uint32)
        before = rng.randint(0, 1000, B).astype(np.uint32)
        for masked in (False, True):
            mask = (rng.randint(0, 4, n) != 0) if masked else None
            tgt = to_dev(before, "u32")
            out = dr.scatter_inc(tgt, to_dev(index, "u32"),
                                 active=torch.from_numpy(mask).cuda() if masked else None)
            check_scatter_inc(before, to_np(tgt, "u32"), index, mask, to_np(out, "u32"), (B, n, masked))
            # the oracle's serial order is one valid answer with the same counters
            exp_tgt, exp_out = capi.scatter_inc(before, index, None if mask is None else mask.astype(np.uint8))
            assert np.array_equal(exp_tgt, to_np(tgt, "u32"))
            assert np.array_equal(np.sort(exp_out), np.sort(to_np(out, "u32")))


@pytest.mark.parametrize("n", [5, 2048, 2049, (1 << 22) + 77, 1184 * 8192 + 8192 + 77])
def test_scatter_inc_queue_form(n):
    """dr.scatter_inc(counter, 0, active): the slots of the active elements are a permutation of
    start .. start + count - 1 (tests/test_while_loop.py:539 uses it as a queue allocator). The largest
    size makes the CTAs of a 148-SM grid (8 per SM, 8192-element tiles) walk more than one tile and ends
    in a partial tile: without a mask array the activity bits of that tile must still be recomputed."""
    mask = (capi.fmix32(n, xor=5) & 1) != 0
    for m, misalign in ((None, 0), (mask, 0), (mask, 1), (None, 3)):     # misaligned: element-wise loads / stores
        tgt = to_dev(np.array([7, 99], np.uint32), "u32")
        d_m = None
        if m is not None:
            d_m = torch.zeros(n + misalign, dtype=torch.bool, device="cuda")[misalign:]
            d_m.copy_(torch.from_numpy(m))
        d_out = torch.full((n + misalign,), -1, dtype=torch.int32, device="cuda")[misalign:]
        out = to_np(dr.scatter_inc(tgt, None, active=d_m, size=n, out=d_out), "u32")
        act = np.ones(n, bool) if m is None else m
        cnt = int(act.sum())
        assert to_np(tgt, "u32").tolist() == [7 + cnt, 99]
        assert np.array_equal(np.sort(out[act]), np.arange(7, 7 + cnt, dtype=np.uint32))
        assert np.all(out[~act] == 0)


@pytest.mark.parametrize("B", [7, 4096])
def test_scatter_inc_ctas_walk_several_tiles(B):
    """Sizes at which a CTA of the 148-SM grid handles more than one tile / round (shared-memory counters
    re-zeroed, bases re-fetched, next-tile prefetch) and the last one is partial"""
    n = 1184 * 2048 + 2048 + 5
    index = capi.fmix32(n, xor=21) % np.uint32(B)
    mask = (capi.fmix32(n, xor=22) % np.uint32(5)) != 0
    before = (np.arange(B, dtype=np.uint32) * 11) % np.uint32(1000)
    for m in (None, mask):
        tgt = to_dev(before, "u32")
        out = dr.scatter_inc(tgt, to_dev(index, "u32"), active=None if m is None else torch.from_numpy(m).cuda())
        check_scatter_inc(before, to_np(tgt, "u32"), index, m, to_np(out, "u32"), (B, m is not None))


def test_scatter_inc_coherent_and_skewed_warps():
    """warps whose lanes agree (one shared-memory atomic), warps with two values, one hot counter"""
    n, B = 1 << 16, 64
    i = np.arange(n, dtype=np.uint32)
    for index in ((i // 32) % B, (i // 16) % B, np.where(i % 5 == 0, i % B, 3).astype(np.uint32)):
        index = index.astype(np.uint32)
        before = np.zeros(B, np.uint32)
        tgt = to_dev(before, "u32")
        out = dr.scatter_inc(tgt, to_dev(index, "u32"))
        check_scatter_inc(before, to_np(tgt, "u32"), index, None, to_np(out, "u32"), "pattern")


def test_scatter_inc_errors_and_out_of_range():
    with pytest.raises(RuntimeError, match="32-bit"):
        dr.scatter_inc(torch.zeros(4, device="cuda"), torch.zeros(4, dtype=torch.int32, device="cuda"))
    # indices past the counter array are ignored (undefined behaviour in the reference)
    for B in (8, 4096):
        index = np.array([0, B, 1, 0xFFFFFFFF, 0], np.uint32)
        tgt = to_dev(np.zeros(B, np.uint32), "u32")
        out = to_np(dr.scatter_inc(tgt, to_dev(index, "u32")), "u32")
        assert to_np(tgt, "u32")[:2].tolist() == [2, 1] and out[1] == 0 and out[3] == 0
        assert sorted(out[[0, 4]].tolist()) == [0, 1] and out[2] == 0


# --------------------------------------------------------------------------- the reference's CUDA JIT beside it
@pytest.fixture(scope="module")
def L():
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    lib = ref.lib(cuda=True, llvm=False)
    if not ref.has_backend(ref.CUDA) or not hasattr(lib, "ref_scatter_packet"):
        pytest.skip("reference CUDA backend (or the packet shims) not available")
    return lib


@pytest.mark.parametrize("vt,op,count", [("f32", "add", 4), ("f32", "add", 2), ("u32", "add", 4), ("f16", "add", 8),
                                         ("f16", "max", 4), ("f32", "min", 4), ("i32", "max", 2), ("u64", "or", 2),
                                         ("f64", "add", 4)])
@pytest.mark.parametrize("mode", [ReduceMode.Direct, ReduceMode.Local])
def test_packet_vs_reference_cuda(L, vt, op, count, mode):
    n, bins = 200_003, 777
    idx, vals = packet_inputs(vt, n, count, bins)
    mask = (capi.fmix32(n, xor=4) & 3) != 0
    d_vals = [to_dev(v, vt) for v in vals]; d_idx = to_dev(idx, "u32"); d_mask = torch.from_numpy(mask).cuda()
    init = make_input(vt, bins * count)
    if vt == "f16":
        init = (capi.fmix32(bins * count, xor=6) % np.uint32(16)).astype(np.float16)
    d_exp = to_dev(init, vt); d_got = to_dev(init, vt)
    torch.cuda.synchronize()
    ptrs = (vp * count)(*[v.data_ptr() for v in d_vals])
    assert L.ref_scatter_packet(ref.CUDA, capi.VT[vt], capi.OP[op], int(mode), vp(d_exp.data_ptr()), bins, ptrs, count,
                                vp(d_idx.data_ptr()), vp(d_mask.data_ptr()), n) == 0
    L.ref_sync()
    dr.scatter_reduce(OPS[op], d_got, d_vals, d_idx, active=d_mask, mode=mode, vt=VT[vt])
    got, exp = to_np(d_got, vt), to_np(d_exp, vt)
    if op == "add" and vt in ("f32", "f64"):
        assert_close(got, exp, vt, n // bins, f"packet {vt} {op} vs reference CUDA")
    else:
        assert np.array_equal(got, exp)


@pytest.mark.parametrize("B", [1, 7, 3000])
def test_scatter_inc_vs_reference_cuda(L, B):
    """Counters bit for bit; slots per counter as sets (both sides hand them out in unspecified order)"""
    n = 150_001
    index = capi.fmix32(n, xor=11) % np.uint32(B)
    mask = (capi.fmix32(n, xor=12) & 3) != 0
    before = np.arange(B, dtype=np.uint32) * 3
    d_idx = to_dev(index, "u32"); d_mask = torch.from_numpy(mask).cuda()
    d_t_ref = to_dev(before, "u32"); d_out_ref = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    assert L.ref_scatter_inc(ref.CUDA, vp(d_t_ref.data_ptr()), B, vp(d_idx.data_ptr()), vp(d_mask.data_ptr()), n,
                             vp(d_out_ref.data_ptr())) == 0
    L.ref_sync()
    d_t = to_dev(before, "u32")
    out = to_np(dr.scatter_inc(d_t, d_idx, active=d_mask), "u32")
    assert torch.equal(d_t, d_t_ref)
    ref_out = to_np(d_out_ref, "u32")
    check_scatter_inc(before, to_np(d_t_ref, "u32"), index, mask, ref_out, "reference")   # the property holds for the reference
    check_scatter_inc(before, to_np(d_t, "u32"), index, mask, out, "b200")
    key = index[mask].astype(np.uint64) << np.uint64(32)
    assert np.array_equal(np.sort(key | out[mask]), np.sort(key | ref_out[mask]))
