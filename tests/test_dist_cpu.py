"""Host-side logic of the sharded path under gloo, world_size 2, on CPU.

drjit_b200.dist.Sharded takes the object that provides the shard-local primitives; here an
oracle-backed CPU stand-in is injected (tests only) so that shard boundaries, the carry
computation of the distributed scan, count/offset exchange and histogram combination are
exercised end to end against the single-array oracle result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import capi  # noqa: E402

_OPN = {1: "add", 2: "mul", 3: "min", 4: "max", 5: "and", 6: "or"}
_VTN = {torch.int32: "u32", torch.float32: "f32", torch.uint8: "u8", torch.float64: "f64"}


class OracleLocal:
    """CPU stand-in for drjit_b200.ops (same call signatures, oracle arithmetic)."""

    @staticmethod
    def _np(x):
        vt = _VTN[x.dtype]
        return vt, x.numpy().view(capi.NP[vt])

    @staticmethod
    def _t(a, like):
        return torch.from_numpy(np.ascontiguousarray(a).view(like.numpy().dtype))

    def block_reduce(self, op, x, block_size, vt=None, out=None):
        t, a = self._np(x)
        return self._t(capi.block_reduce(t, _OPN[int(op)], a, block_size, acc64=True), x)

    def all(self, mask):    # noqa: A003
        return bool(np.all(mask.numpy() != 0))

    def any(self, mask):    # noqa: A003
        return bool(np.any(mask.numpy() != 0))

    def dot(self, a, b):
        return torch.tensor([float(capi.reduce_dot("f32", a.numpy(), b.numpy(), acc64=True))], dtype=torch.float32)

    def block_prefix_reduce(self, op, x, block_size, exclusive=True, reverse=False, vt=None, out=None):
        t, a = self._np(x)
        return self._t(capi.block_prefix_reduce(t, _OPN[int(op)], a, block_size, exclusive, reverse), x)

    def prefix_reduce_carry(self, op, x, exclusive=True, reverse=False, carry_in=None, total_out=None, vt=None, out=None):
        t, a = self._np(x)
        res = capi.block_prefix_reduce(t, _OPN[int(op)], a, a.size, exclusive, reverse)
        if carry_in is not None:
            res = (res + self._np(carry_in)[1][0]).astype(res.dtype)   # Add only (what the tests use)
        if total_out is not None:
            tot = capi.block_reduce(t, _OPN[int(op)], a, a.size)
            if carry_in is not None:
                tot = (tot + self._np(carry_in)[1][0]).astype(tot.dtype)
            total_out.copy_(self._t(tot, x))
        return self._t(res, x)

    def compress_async(self, mask, index_base=0, out=None, count=None):
        idx = capi.compress(mask.numpy()) + np.uint32(index_base)
        o = torch.zeros(mask.numel(), dtype=torch.int32)
        o[:idx.size] = torch.from_numpy(idx.view(np.int32))
        return o, torch.tensor([idx.size], dtype=torch.int32)

    def mkperm_sharded(self, values, bucket_count, index_base=0, perm=None, hist=None):
        keys = values.numpy().view(np.uint32)
        p, _, _ = capi.block_mkperm(keys, keys.size, bucket_count)
        h = np.bincount(keys, minlength=bucket_count).astype(np.int32)
        return torch.from_numpy((p + np.uint32(index_base)).view(np.int32)), torch.from_numpy(h)

    def scatter_reduce(self, op, target, value, index, active=None, mode=0, vt=None):
        res = capi.scatter_reduce("f32", "add", target.numpy(), value.numpy(), index.numpy().view(np.uint32), acc64=True)
        target.copy_(torch.from_numpy(res))
        return target


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from drjit_b200.dist import Sharded
        from drjit_b200.ops import ReduceOp
        sh = Sharded(rank=rank, world=world, group=dist.group.WORLD, local=OracleLocal())

        lo, hi = sh.shard_range(n, align=256)
        u = capi.fmix32(n)
        ut = torch.from_numpy(u[lo:hi].view(np.int32).copy())

        # sum / min / max
        for op, name in ((ReduceOp.Add, "add"), (ReduceOp.Min, "min"), (ReduceOp.Max, "max")):
            got = sh.reduce(op, ut).numpy().view(np.uint32)[0]
            assert got == capi.block_reduce("u32", name, u, n)[0], name

        # block_reduce: shards cut at block boundaries reduce independently, outputs concatenate
        got = sh.block_reduce(ReduceOp.Add, ut, 256).numpy().view(np.uint32)
        exp = capi.block_reduce("u32", "add", u, 256)
        assert np.array_equal(got, exp[lo // 256: lo // 256 + got.size]) and got.size == -(-(hi - lo) // 256)
        pieces = [None] * world
        dist.all_gather_object(pieces, got.size)
        assert sum(pieces) == exp.size

        # all / any: every rank sees the global answer, whichever shard holds the deciding byte
        for pos in (0, n - 1, n // 2):
            ones = np.ones(n, np.uint8); ones[pos] = 0
            zeros = np.zeros(n, np.uint8); zeros[pos] = 1
            assert sh.all(torch.from_numpy(ones[lo:hi].copy())) is False
            assert sh.any(torch.from_numpy(zeros[lo:hi].copy())) is True
        assert sh.all(torch.ones(hi - lo, dtype=torch.uint8)) is True
        assert sh.any(torch.zeros(hi - lo, dtype=torch.uint8)) is False

        # distributed exclusive prefix sum == slice of the single-array oracle scan (bit-exact)
        got = sh.prefix_sum(ut).numpy().view(np.uint32)
        exp = capi.block_prefix_reduce("u32", "add", u, n, True, False)[lo:hi]
        assert np.array_equal(got, exp)

        # shard-offset form: offset (+) local == the same slice
        local, off = sh.prefix_reduce_offsets(ReduceOp.Add, ut)
        assert np.array_equal((local.numpy().view(np.uint32) + off.numpy().view(np.uint32)[0]).astype(np.uint32), exp)

        # fold of one scalar per rank (all ranks / only the lower ones)
        mine = torch.tensor([rank + 5], dtype=torch.int32)
        assert int(sh.fold_scalar(ReduceOp.Add, mine, torch.zeros(1, dtype=torch.int32))[0]) == sum(r + 5 for r in range(world))
        assert int(sh.fold_scalar(ReduceOp.Add, mine, torch.zeros(1, dtype=torch.int32), lower=True)[0]) == sum(r + 5 for r in range(rank))

        # compress: global indices, rank-order concatenation == oracle list
        m = capi.mask_u8(n, 128)
        out, counts = sh.compress(torch.from_numpy(m[lo:hi].copy()), lo)
        exp_all = capi.compress(m)
        start = sum(counts[:rank])
        assert sum(counts) == exp_all.size
        assert np.array_equal(out[:counts[rank]].numpy().view(np.uint32), exp_all[start:start + counts[rank]])

        # mkperm: global histogram + rank-major stable order == oracle permutation
        B = 37
        keys = capi.fmix32(n) % np.uint32(B)
        res = sh.mkperm(torch.from_numpy(keys[lo:hi].view(np.int32).copy()), B, lo)
        perm, hist = res.perm, res.hist
        ghist = np.bincount(keys, minlength=B)
        exp_perm, exp_off, exp_unique = capi.block_mkperm(keys, n, B)
        # the table of non-empty buckets describes the GLOBAL array: identical to the single-array oracle table
        assert res.table.shape[0] == exp_unique
        assert np.array_equal(res.table.numpy().astype(np.uint32).reshape(-1), exp_off[:4 * exp_unique])
        local_start = np.cumsum(hist.numpy()) - hist.numpy()
        rank_base = res.rank_base.numpy().view(np.uint32)
        for b in range(B):      # rank r's slice of bucket b sits at rank_base[b] of the stable global permutation
            mine = perm.numpy().view(np.uint32)[local_start[b]:local_start[b] + int(hist[b])]
            assert np.array_equal(mine, exp_perm[rank_base[b]: rank_base[b] + int(hist[b])])
        assert int(hist.sum()) == hi - lo and ghist.sum() == n

        # scatter-add with all-reduced bins; dot
        f = capi.unit_f32(n)
        idx = capi.fmix32(n, xor=0x85EBCA6B, mask=63)
        bins = sh.scatter_add(torch.zeros(64), torch.from_numpy(f[lo:hi].copy()),
                              torch.from_numpy(idx[lo:hi].view(np.int32).copy()))
        exp = capi.scatter_reduce("f32", "add", np.zeros(64, np.float32), f, idx, acc64=True)
        assert np.allclose(bins.numpy(), exp, rtol=1e-5)
        d = sh.dot(torch.from_numpy(f[lo:hi].copy()), torch.from_numpy(f[lo:hi].copy()))
        assert abs(float(d[0]) - float(np.dot(f.astype(np.float64), f))) < 1e-4 * n
    finally:
        dist.destroy_process_group()


def _worker_tiny(rank, world, port):
    """n < world: the trailing shard is empty and must contribute the identity (no size-mismatched
    collective, no uninitialised total)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from drjit_b200.dist import Sharded
        from drjit_b200.ops import ReduceOp, VarType
        sh = Sharded(rank=rank, world=world, group=dist.group.WORLD, local=OracleLocal())
        n = 1
        lo, hi = sh.shard_range(n)
        assert (hi - lo) == (1 if rank == 0 else 0)
        u = np.array([41], np.uint32)
        ut = torch.from_numpy(u[lo:hi].view(np.int32).copy())
        for op, exp in ((ReduceOp.Add, 41), (ReduceOp.Min, 41), (ReduceOp.Max, 41), (ReduceOp.Mul, 41)):
            assert sh.reduce(op, ut, vt=VarType.UInt32).numpy().view(np.uint32)[0] == exp, op
        got = sh.prefix_sum(ut).numpy().view(np.uint32)
        assert got.size == hi - lo and (got.size == 0 or got[0] == 0)
        local, off = sh.prefix_reduce_offsets(ReduceOp.Add, ut)
        assert int(off.numpy().view(np.uint32)[0]) == (0 if rank == 0 else 41)
        assert sh.all(torch.ones(hi - lo, dtype=torch.uint8)) is True
        assert sh.any(torch.zeros(hi - lo, dtype=torch.uint8)) is False
        out, counts = sh.compress(torch.ones(hi - lo, dtype=torch.uint8), lo)
        assert counts == [1, 0]
        res = sh.mkperm(torch.zeros(hi - lo, dtype=torch.int32), 4, lo)
        assert res.table.tolist() == [[0, 0, 1, 0]] and int(res.hist.sum()) == hi - lo
        f = torch.full((hi - lo,), 2.0)
        assert float(sh.dot(f, f)[0]) == 4.0
        bins = sh.scatter_add(torch.zeros(4), f, torch.zeros(hi - lo, dtype=torch.int32))
        assert bins.tolist() == [2.0, 0.0, 0.0, 0.0]
    finally:
        dist.destroy_process_group()


def test_sharded_empty_trailing_shard_gloo_world2():
    mp.spawn(_worker_tiny, args=(2, 29400 + (os.getpid() % 90)), nprocs=2, join=True)


@pytest.mark.parametrize("n", [100_003, 4096])
def test_sharded_primitives_gloo_world2(n):
    port = 29500 + (os.getpid() % 500) + (1 if n == 4096 else 0)
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)


def test_shard_bounds():
    from drjit_b200.dist import shard_bounds
    for n in (0, 1, 1000, 1 << 20, (1 << 20) + 7):
        for world in (1, 2, 4, 8):
            for align in (1, 256, 1024):
                b = shard_bounds(n, world, align)
                assert b[0] == 0 and b[-1] == n and len(b) == world + 1
                assert all(x <= y for x, y in zip(b, b[1:]))
                assert all(x % align == 0 or x == n for x in b)
