"""Second oracle on the GPU box: the UNMODIFIED reference's own CUDA backend (oracle/_ref, kernels
PTX-JITed by the driver) run on the same device buffers. Integer outputs must agree bit for bit
(scan, compress); block_mkperm must produce the same bucket table rows and the same elements per
bucket (the reference's CUDA variant is not stable and appends table rows through an atomic
counter, resources/mkperm.cuh:309-317, so order inside a bucket / of the rows is not compared).
Skipped when oracle/_ref is absent or its CUDA backend does not initialise."""
import ctypes

import numpy as np
import pytest
import torch

import drjit_b200 as dr
from drjit_b200 import ReduceOp, VarType, ops
from oracle import ref
from oracle.capi import OP, VT

pytestmark = pytest.mark.gpu
vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def L():
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    lib = ref.lib(cuda=True, llvm=False)
    if not ref.has_backend(ref.CUDA):
        pytest.skip("reference CUDA backend did not initialise")
    return lib


@pytest.mark.parametrize("n", [1000, (1 << 20) + 5, (1 << 23) + 4])
def test_prefix_sum_vs_reference_cuda(L, n):
    x = torch.empty(n, dtype=torch.int32, device="cuda"); ops.fill_fmix32(x, 0)
    for bs in (n, 1000, 256):
        for ex, rev in ((1, 0), (0, 0), (1, 1)):
            exp = torch.empty_like(x)
            torch.cuda.synchronize()
            assert L.ref_block_prefix_reduce(ref.CUDA, VT["u32"], OP["add"], n, bs, ex, rev, vp(x.data_ptr()), vp(exp.data_ptr())) == 0
            L.ref_sync()
            got = ops.block_prefix_reduce(ReduceOp.Add, x, bs, bool(ex), bool(rev), vt=VarType.UInt32)
            assert torch.equal(got, exp), (n, bs, ex, rev)


@pytest.mark.parametrize("n", [4097, (1 << 22) + 3])
def test_compress_vs_reference_cuda(L, n):
    buf = torch.zeros(n + 4096, dtype=torch.uint8, device="cuda")    # the reference zero-pads to a multiple of 2048
    m = buf[:n]; ops.fill_fmix32(m, 2, and_=77)
    exp = torch.empty(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    cnt = L.ref_compress(ref.CUDA, vp(m.data_ptr()), n, vp(exp.data_ptr()))
    L.ref_sync()
    got = dr.compress(m)
    assert cnt == got.numel() and torch.equal(exp[:cnt], got.view(torch.int32))


def test_mkperm_vs_reference_cuda(L):
    n, B = (1 << 22) + 11, 1500
    keys = torch.empty(n, dtype=torch.int32, device="cuda"); ops.fill_fmix32(keys, 0)
    keys = (keys.to(torch.int64) & 0xFFFFFFFF).remainder(B - 3).to(torch.int32)      # three buckets stay empty
    perm_r = torch.empty_like(keys)
    off = L.ref_malloc(ref.CUDA, 4 * (4 * B + 1), 1)
    torch.cuda.synchronize()
    uniq = L.ref_block_mkperm(ref.CUDA, vp(keys.data_ptr()), n, n, B, vp(perm_r.data_ptr()), vp(off))
    L.ref_sync()
    tab_r = np.ctypeslib.as_array((ctypes.c_uint32 * (4 * B + 1)).from_address(off)).copy()
    L.ref_free(off)
    perm_o, table = dr.block_mkperm(keys, n, B)
    torch.cuda.synchronize()
    tab_o = table.numpy().astype(np.uint32).reshape(-1, 4)
    tab_r = tab_r[:4 * uniq].reshape(-1, 4)
    tab_r = tab_r[np.argsort(tab_r[:, 0], kind="stable")]
    assert uniq == tab_o.shape[0] == B - 3 and np.array_equal(tab_o[:, :3], tab_r[:, :3])
    k64 = keys.to(torch.int64)
    pr, po = perm_r.to(torch.int64), perm_o.view(torch.int32).to(torch.int64)
    assert torch.equal(torch.sort(k64[pr] * n + pr).values, torch.sort(k64[po] * n + po).values)
    assert bool(torch.all(k64[po][1:] >= k64[po][:-1]))


def test_baseline_sizes_vs_reference_cuda_full_arrays(L):
    """The BASELINE configurations checked over the WHOLE arrays, on the device, against the reference's
    own CUDA kernels: exclusive u32 prefix sum of 2^30 elements and the index list of a 2^30-byte mask bit
    for bit; block_mkperm of 2^26 IDs into 4096 buckets by table rows and per-bucket contents."""
    n = 1 << 30
    x = torch.empty(n, dtype=torch.int32, device="cuda"); ops.fill_fmix32(x, 0)
    exp = torch.empty_like(x)
    torch.cuda.synchronize()
    assert L.ref_block_prefix_reduce(ref.CUDA, VT["u32"], OP["add"], n, n, 1, 0, vp(x.data_ptr()), vp(exp.data_ptr())) == 0
    L.ref_sync()
    got = ops.block_prefix_reduce(ReduceOp.Add, x, n, True, False, vt=VarType.UInt32)
    assert torch.equal(got, exp)
    del x, got

    buf = torch.zeros(n + 4096, dtype=torch.uint8, device="cuda")    # the reference zero-pads to a multiple of 2048
    m = buf[:n]; ops.fill_fmix32(m, 2, and_=128)
    torch.cuda.synchronize()
    cnt = L.ref_compress(ref.CUDA, vp(m.data_ptr()), n, vp(exp.data_ptr()))
    L.ref_sync()
    got = dr.compress(m)
    assert cnt == got.numel() and torch.equal(exp[:cnt], got.view(torch.int32))
    del buf, m, got, exp

    n, B = 1 << 26, 4096
    keys = torch.empty(n, dtype=torch.int32, device="cuda"); ops.fill_fmix32(keys, 0, and_=B - 1)
    perm_r = torch.empty_like(keys)
    off = L.ref_malloc(ref.CUDA, 4 * (4 * B + 1), 1)
    torch.cuda.synchronize()
    uniq = L.ref_block_mkperm(ref.CUDA, vp(keys.data_ptr()), n, n, B, vp(perm_r.data_ptr()), vp(off))
    L.ref_sync()
    tab_r = np.ctypeslib.as_array((ctypes.c_uint32 * (4 * B + 1)).from_address(off)).copy()
    L.ref_free(off)
    perm_o, table = dr.block_mkperm(keys, n, B)
    torch.cuda.synchronize()
    tab_o = table.numpy().astype(np.uint32).reshape(-1, 4)
    tab_r = tab_r[:4 * uniq].reshape(-1, 4)
    tab_r = tab_r[np.argsort(tab_r[:, 0], kind="stable")]
    assert uniq == tab_o.shape[0] and np.array_equal(tab_o[:, :3], tab_r[:, :3])
    k64 = keys.to(torch.int64)
    for perm in (perm_r, perm_o.view(torch.int32)):
        p = perm.to(torch.int64)
        kp = k64[p]
        assert bool(torch.all(kp[1:] >= kp[:-1]))                              # grouped by bucket
        assert torch.equal(torch.sort(p).values, torch.arange(n, device="cuda"))    # a permutation
