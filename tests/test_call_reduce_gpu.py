"""jit_var_call_reduce as one call (ext/drjit-core/src/call.cpp:1268-1389):
block_mkperm of the callable IDs + the bucket table in the dispatcher's order (decreasing size, sorted
on the device) + argument arrays permuted by the scatter pass itself. Checked against the oracle's
block_mkperm (bucket contents, table rows) and numpy gathers."""
import numpy as np
import pytest
import torch

import drjit_b200 as dr
from oracle import capi

pytestmark = pytest.mark.gpu


def _keys(n, buckets, skew):
    k = capi.fmix32(n) % np.uint32(buckets)
    if skew:                                    # min of two draws: sizes differ a lot between buckets
        k = np.minimum(k, capi.fmix32(n, xor=0x9E3779B9) % np.uint32(buckets))
    return k.astype(np.uint32)


@pytest.mark.parametrize("n,buckets", [(1, 4), (1000, 37), (70_001, 4096), ((1 << 20) + 77, 4096),
                                       ((1 << 20) + 77, 256), ((1 << 22) + 3, 1000), (300_000, 20_000)])
@pytest.mark.parametrize("skew", [False, True])
def test_call_reduce_table_and_payloads(n, buckets, skew):
    keys = _keys(n, buckets, skew)
    a = capi.unit_f32(n)
    b = capi.fmix32(n, xor=7)
    perm, table, (pa, pb) = dr.call_reduce(torch.from_numpy(keys.view(np.int32)).cuda(), buckets,
                                           [torch.from_numpy(a).cuda(), torch.from_numpy(b.view(np.int32)).cuda()])
    torch.cuda.synchronize()
    p = perm.cpu().numpy().view(np.uint32)
    eperm, eoff, eunique = capi.block_mkperm(keys, n, buckets)
    rows = table.numpy()
    # same rows as the reference table, ordered by decreasing size (ties: ascending id)
    exp_rows = eoff[:4 * eunique].reshape(-1, 4).astype(np.int64)
    order = np.lexsort((exp_rows[:, 0], -exp_rows[:, 2]))
    assert rows.shape[0] == eunique and np.array_equal(rows, exp_rows[order])
    assert np.all(np.diff(rows[:, 2]) <= 0)
    # bucket contents: equal as sets (the permutation is stable only up to 1816 buckets, jit.h:2404-2406)
    stable = buckets * 4 * 32 <= 227 * 1024 or n < (1 << 18)
    if stable:
        assert np.array_equal(p, eperm)
    else:
        assert np.array_equal(np.sort(p), np.arange(n, dtype=np.uint32)) and np.all(np.diff(keys[p].astype(np.int64)) >= 0)
    # payloads travelled with the permutation
    assert np.array_equal(pa.cpu().numpy(), a[p])
    assert np.array_equal(pb.cpu().numpy().view(np.uint32), b[p])


def test_call_reduce_without_payloads_matches_block_mkperm():
    n, buckets = (1 << 20) + 5, 512
    keys = torch.from_numpy(_keys(n, buckets, True).view(np.int32)).cuda()
    perm, table, outs = dr.call_reduce(keys, buckets)
    perm2, table2 = dr.block_mkperm(keys, n, buckets)
    torch.cuda.synchronize()
    assert outs == [] and torch.equal(perm, perm2)
    t2 = table2.numpy()
    assert np.array_equal(table.numpy(), t2[np.lexsort((t2[:, 0], -t2[:, 2]))])


@pytest.mark.parametrize("buckets,npay", [(4096, 1), (4096, 4), (8000, 2)])
def test_call_reduce_staged_payload_tiles(buckets, npay):
    """Inputs large enough for the payload-staging tiles (24 Ki / 20 Ki keys, the payload tile in the
    other half of shared memory): ragged last tile, one payload array deliberately not 16-byte
    aligned (element-wise tile load), 1 .. 4 arrays, skewed IDs."""
    n = 148 * 2 * 1024 * 24 + 12_345
    keys = _keys(n, buckets, True)
    pays = [capi.fmix32(n + 1, xor=11 + k) for k in range(npay)]
    dev = [torch.from_numpy(x.view(np.int32)).cuda() for x in pays]
    ins = [d[1:] if k == 0 else d[:n] for k, d in enumerate(dev)]         # payload 0 starts 4 bytes off alignment
    perm, table, outs = dr.call_reduce(torch.from_numpy(keys.view(np.int32)).cuda(), buckets, ins)
    torch.cuda.synchronize()
    p = perm.cpu().numpy().view(np.uint32)
    assert np.array_equal(np.sort(p), np.arange(n, dtype=np.uint32)) and np.all(np.diff(keys[p].astype(np.int64)) >= 0)
    hist = np.bincount(keys, minlength=buckets)
    rows = table.numpy()
    assert rows.shape[0] == int((hist > 0).sum()) and np.all(np.diff(rows[:, 2]) <= 0)
    assert np.array_equal(rows[:, 2], hist[rows[:, 0]])
    for k, o in enumerate(outs):
        src = pays[k][1:] if k == 0 else pays[k][:n]
        assert np.array_equal(o.cpu().numpy().view(np.uint32), src[p]), k
