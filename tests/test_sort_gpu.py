"""dr.sort / dr.argsort against the reference's definition: a stable sort
by the order-preserving unsigned image of the keys (`_to_ordinal_32/_64`, drjit/__init__.py:1483-1520,
`_radix_sort` :1698-1772). Small and ragged sizes are checked bit for bit against numpy's stable
argsort of that image (the oracle here is ten lines of numpy: the reference's own formula); the
BASELINE-sized run (2^26 keys) against torch.sort(stable=True) on the device."""
import numpy as np
import pytest
import torch

import drjit_b200 as dr
from drjit_b200 import VarType
from oracle import capi

pytestmark = pytest.mark.gpu


def ordinal(a):
    """numpy restatement of _to_ordinal_32 / _to_ordinal_64"""
    bits = a.dtype.itemsize * 8
    U = np.uint32 if bits == 32 else np.uint64
    u = a.view(U)
    top = U(1) << U(bits - 1)
    if a.dtype.kind == "f":
        mask = (U(0) - (u >> U(bits - 1))).astype(U)
        return u ^ (mask | top)
    if a.dtype.kind == "i":
        return u ^ top
    return u


def expected(a, descending):
    o = ordinal(a)
    if descending:
        o = ~o
    perm = np.argsort(o, kind="stable").astype(np.uint32)
    return a[perm], perm


def make(dtype, n, seed):
    raw = capi.fmix32(2 * n, start=seed)
    if dtype in (np.uint32, np.int32):
        a = raw[:n].view(dtype).copy()
        a[::7] = a[0] if n else 0            # duplicates: stability matters
    elif dtype == np.float32:
        a = (raw[:n].astype(np.float64) / 2 ** 31 - 1.0).astype(np.float32) * np.float32(1e3)
        a[::5] = np.float32(0.25)
        if n > 8:
            a[1] = -0.0; a[2] = 0.0; a[3] = np.inf; a[4] = -np.inf; a[5] = np.nan
            a[6] = np.float32(np.nan) * np.float32(-1)
    else:
        a64 = raw.view(np.uint64)[:n].copy()
        if dtype == np.float64:
            a = (a64.astype(np.float64) / 2 ** 63 - 1.0) * 1e6
            a[::5] = 0.5
        else:
            a = a64.view(dtype).copy()
            a[::7] = a[0] if n else 0
    return np.ascontiguousarray(a, dtype)


_T = {np.uint32: (torch.int32, VarType.UInt32), np.int32: (torch.int32, VarType.Int32),
      np.float32: (torch.float32, VarType.Float32), np.uint64: (torch.int64, VarType.UInt64),
      np.int64: (torch.int64, VarType.Int64), np.float64: (torch.float64, VarType.Float64)}


def to_dev(a):
    tdt, vt = _T[a.dtype.type]
    signed = {np.uint32: np.int32, np.uint64: np.int64}.get(a.dtype.type, a.dtype.type)
    return torch.from_numpy(a.view(signed)).cuda(), vt


@pytest.mark.parametrize("dtype", [np.uint32, np.int32, np.float32, np.uint64, np.int64, np.float64])
@pytest.mark.parametrize("n", [1, 2, 33, 1000, 8192, 8193, (1 << 20) + 77])
def test_sort_argsort_bit_exact(dtype, n):
    a = make(dtype, n, seed=n)
    t, vt = to_dev(a)
    for descending in (False, True):
        ev, ep = expected(a, descending)
        values, index = dr.sort_with_indices(t, descending, vt=vt)
        got_v = values.cpu().numpy().view(dtype)
        got_p = index.cpu().numpy().view(np.uint32)
        assert np.array_equal(got_p, ep), (dtype, n, descending)
        assert got_v.tobytes() == ev.tobytes(), (dtype, n, descending)      # (bytes: NaNs compare equal)
        # the single-output forms run the same passes without the other array
        bits = torch.int32 if a.dtype.itemsize == 4 else torch.int64     # (bit patterns: NaN == NaN)
        assert torch.equal(dr.sort(t, descending, vt=vt).view(bits), values.view(bits))
        assert torch.equal(dr.argsort(t, descending, vt=vt), index)
    assert torch.equal(t.view(bits), to_dev(a)[0].view(bits))          # the input is never written


def test_sort_unaligned_input():
    n = 100_003
    a = make(np.uint32, n + 3, seed=5)
    buf, vt = to_dev(a)
    t = buf[3:]                                   # 12-byte offset: no vector loads
    ev, ep = expected(a[3:], False)
    values, index = dr.sort_with_indices(t, False, vt=vt)
    assert np.array_equal(index.cpu().numpy().view(np.uint32), ep)
    assert np.array_equal(values.cpu().numpy().view(np.uint32), ev)


@pytest.mark.parametrize("dtype", [torch.int32, torch.float32])
def test_sort_2_26_vs_torch_stable(dtype):
    n = 1 << 26
    raw = torch.empty(n, dtype=torch.int32, device="cuda")
    dr.ops.fill_fmix32(raw, 0, and_=0x000FFFFF if dtype == torch.int32 else 0xFFFFFFFF)   # int: many duplicates
    if dtype == torch.float32:
        t = (raw.to(torch.float32) * 1e-3).contiguous()
    else:
        t = raw - (1 << 19)
    values, index = dr.sort_with_indices(t)
    ev, ep = torch.sort(t, stable=True)
    assert torch.equal(values, ev)
    assert torch.equal(index.to(torch.int64), ep)
    del ev, ep
    dv = dr.sort(t, descending=True)
    assert torch.equal(dv, torch.sort(t, descending=True, stable=True)[0])


def test_sort_rejects_unsupported_type():
    with pytest.raises(RuntimeError):
        dr.sort(torch.zeros(16, dtype=torch.float16, device="cuda"))
