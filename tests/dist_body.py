"""Checks of the sharded primitives against the single-array oracle, shared by the in-process
(threads, peer access) and the multi-process (torchrun-style, CUDA IPC / NCCL) GPU tests.
`sh` is a drjit_b200.dist.Sharded for rank `rank` of `world`, `dev` its device."""
import numpy as np
import torch

from oracle import capi


def check_rank(sh, rank, world, dev, n, buckets=4096, bins=1 << 12):
    from drjit_b200.ops import ReduceOp, VarType

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    lo, hi = sh.shard_range(n, align=256)
    u = capi.fmix32(n)
    ut = up(u[lo:hi].view(np.int32))

    # ---- dr.sum / min / max / prod on u32 (bit-exact, every rank holds the global value)
    for op, name in ((ReduceOp.Add, "add"), (ReduceOp.Min, "min"), (ReduceOp.Max, "max"), (ReduceOp.Mul, "mul")):
        got = sh.reduce(op, ut, vt=VarType.UInt32).cpu().numpy().view(np.uint32)[0]
        assert got == capi.block_reduce("u32", name, u, n)[0], (name, rank)
    # u64 min (8-byte payload) and f32 sum (tolerance 1e-6 * log2 N, north_star)
    u64 = capi.fmix32_u64(n)
    got = sh.reduce(ReduceOp.Min, up(u64[lo:hi].view(np.int64)), vt=VarType.UInt64).cpu().numpy().view(np.uint64)[0]
    assert got == u64.min()
    f = capi.unit_f32(n)
    ft = up(f[lo:hi])
    got = float(sh.reduce(ReduceOp.Add, ft).cpu()[0])
    exp = float(f.astype(np.float64).sum())
    assert abs(got - exp) <= 1e-6 * np.log2(max(n, 2)) * abs(exp)
    # f16 (accumulates in f32) and u8
    h = (capi.fmix32(n) % np.uint32(8)).astype(np.float16)
    got = float(sh.reduce(ReduceOp.Max, up(h[lo:hi])).cpu()[0])
    assert got == float(h.max())

    # ---- block_reduce: shards cut at block boundaries reduce independently
    got = sh.block_reduce(ReduceOp.Add, ut, 256, vt=VarType.UInt32).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, capi.block_reduce("u32", "add", u, 256)[lo // 256: lo // 256 + got.size])

    # ---- dr.all / dr.any
    for pos in (0, n - 1, n // 2):
        ones = np.ones(n, np.uint8); ones[pos] = 0
        zeros = np.zeros(n, np.uint8); zeros[pos] = 1
        assert sh.all(up(ones[lo:hi])) is False
        assert sh.any(up(zeros[lo:hi])) is True
    assert sh.all(torch.ones(hi - lo, dtype=torch.uint8, device=dev)) is True
    assert sh.any(torch.zeros(hi - lo, dtype=torch.uint8, device=dev)) is False

    # ---- dot
    d = float(sh.dot(ft, ft).cpu()[0])
    ref = float(np.dot(f.astype(np.float64), f.astype(np.float64)))
    assert abs(d - ref) <= 1e-6 * np.log2(max(n, 2)) * ref

    # ---- exclusive prefix sum, materialised and in shard-offset form (bit-exact)
    exp = capi.block_prefix_reduce("u32", "add", u, n, True, False)[lo:hi]
    got = sh.prefix_sum(ut, vt=VarType.UInt32).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, exp)
    local, off = sh.prefix_reduce_offsets(ReduceOp.Add, ut, vt=VarType.UInt32)
    got = (local.cpu().numpy().view(np.uint32) + off.cpu().numpy().view(np.uint32)[0]).astype(np.uint32)
    assert np.array_equal(got, exp)
    # inclusive max scan of u64 (16-byte tile descriptors + 8-byte exchange payload)
    exp = capi.block_prefix_reduce("u64", "max", u64, n, False, False)[lo:hi]
    got = sh.prefix_reduce(ReduceOp.Max, up(u64[lo:hi].view(np.int64)), exclusive=False, vt=VarType.UInt64)
    assert np.array_equal(got.cpu().numpy().view(np.uint64), exp)

    # ---- fold of one scalar per rank (the e2e path folds the chunked scan's shard totals with it)
    mine = torch.tensor([rank + 5], dtype=torch.int32, device=dev)
    got = sh.fold_scalar(ReduceOp.Add, mine, torch.zeros(1, dtype=torch.int32, device=dev), vt=VarType.UInt32)
    assert int(got.cpu()[0]) == sum(r + 5 for r in range(world))
    got = sh.fold_scalar(ReduceOp.Add, mine, torch.zeros(1, dtype=torch.int32, device=dev), lower=True, vt=VarType.UInt32)
    assert int(got.cpu()[0]) == sum(r + 5 for r in range(rank))

    # ---- compress: global indices, rank-order concatenation == oracle list
    m = capi.mask_u8(n, 128)
    out, counts = sh.compress(up(m[lo:hi]), lo)
    exp_all = capi.compress(m)
    start = sum(counts[:rank])
    assert len(counts) == world and sum(counts) == exp_all.size
    assert np.array_equal(out[:counts[rank]].cpu().numpy().view(np.uint32), exp_all[start:start + counts[rank]])

    # ---- mkperm: global bucket table == oracle table; shard slices sit at rank_base of the global order
    for B in (buckets, 37):
        keys = capi.fmix32(n) % np.uint32(B)
        res = sh.mkperm(up(keys[lo:hi].view(np.int32)), B, lo)
        torch.cuda.synchronize(dev)
        exp_perm, exp_off, exp_unique = capi.block_mkperm(keys, n, B)
        assert res.table.shape[0] == exp_unique
        assert np.array_equal(res.table.numpy().astype(np.uint32).reshape(-1), exp_off[:4 * exp_unique])
        hist = res.hist.cpu().numpy().view(np.uint32)
        assert np.array_equal(hist, np.bincount(keys[lo:hi], minlength=B))
        p = res.perm.cpu().numpy().view(np.uint32)
        rank_base = res.rank_base.cpu().numpy().view(np.uint32)
        local_start = np.cumsum(hist) - hist
        stable = B * 4 * 32 <= 227 * 1024 or (hi - lo) < (1 << 18)
        for b in range(B):
            mine = p[local_start[b]:local_start[b] + hist[b]]
            want = exp_perm[rank_base[b]: rank_base[b] + hist[b]]
            if stable:
                assert np.array_equal(mine, want), (B, b)
            else:       # unordered inside a bucket beyond 1816 buckets, like the reference
                assert np.array_equal(np.sort(mine), np.sort(want)), (B, b)

    # ---- scatter-add: bins summed over the ranks
    idx = capi.fmix32(n, xor=0x85EBCA6B, mask=bins - 1)
    got = sh.scatter_add(torch.zeros(bins, device=dev), ft, up(idx[lo:hi].view(np.int32))).cpu().numpy()
    exp = capi.scatter_reduce("f32", "add", np.zeros(bins, np.float32), f, idx, acc64=True)
    assert np.all(np.abs(got - exp) <= 1e-5 * np.maximum(np.abs(exp), 1))
    return got      # (callers compare the bins of all ranks bit for bit)
