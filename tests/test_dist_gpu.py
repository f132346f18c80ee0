"""The sharded path on real GPUs: one process per GPU, NCCL, world_size 2 (skipped on boxes with
fewer than two devices). Same checks as test_dist_cpu.py, but with the CUDA library as the
shard-local implementation and the C oracle as the checker."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import capi

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from drjit_b200.dist import Sharded
        from drjit_b200.ops import ReduceOp, VarType
        sh = Sharded(rank=rank, world=world, group=dist.group.WORLD)
        lo, hi = sh.shard_range(n, align=256)
        u = capi.fmix32(n)
        ut = torch.from_numpy(u[lo:hi].view(np.int32).copy()).to(dev)

        for op, name in ((ReduceOp.Add, "add"), (ReduceOp.Min, "min"), (ReduceOp.Max, "max")):
            got = sh.reduce(op, ut, vt=VarType.UInt32).cpu().numpy().view(np.uint32)[0]
            assert got == capi.block_reduce("u32", name, u, n)[0], name

        got = sh.block_reduce(ReduceOp.Add, ut, 256, vt=VarType.UInt32).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, capi.block_reduce("u32", "add", u, 256)[lo // 256: lo // 256 + got.size])

        ones = np.ones(n, np.uint8); ones[n - 1] = 0
        zeros = np.zeros(n, np.uint8); zeros[0] = 1
        assert sh.all(torch.from_numpy(ones[lo:hi].copy()).to(dev)) is False
        assert sh.any(torch.from_numpy(zeros[lo:hi].copy()).to(dev)) is True
        assert sh.all(torch.ones(hi - lo, dtype=torch.uint8, device=dev)) is True
        assert sh.any(torch.zeros(hi - lo, dtype=torch.uint8, device=dev)) is False

        got = sh.prefix_sum(ut, vt=VarType.UInt32).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, capi.block_prefix_reduce("u32", "add", u, n, True, False)[lo:hi])

        local, off = sh.prefix_reduce_offsets(ReduceOp.Add, ut, vt=VarType.UInt32)
        got = (local.cpu().numpy().view(np.uint32) + off.cpu().numpy().view(np.uint32)[0]).astype(np.uint32)
        assert np.array_equal(got, capi.block_prefix_reduce("u32", "add", u, n, True, False)[lo:hi])

        m = capi.mask_u8(n, 128)
        out, counts = sh.compress(torch.from_numpy(m[lo:hi].copy()).to(dev), lo)
        exp_all = capi.compress(m)
        start = sum(counts[:rank])
        assert sum(counts) == exp_all.size
        assert np.array_equal(out[:counts[rank]].cpu().numpy().view(np.uint32), exp_all[start:start + counts[rank]])

        B = 4096
        keys = capi.fmix32(n, mask=B - 1)
        perm, hist, ghist = sh.mkperm(torch.from_numpy(keys[lo:hi].view(np.int32).copy()).to(dev), B, lo)
        torch.cuda.synchronize()
        assert np.array_equal(ghist.numpy(), np.bincount(keys, minlength=B))
        h = hist.cpu().numpy()
        assert np.array_equal(h, np.bincount(keys[lo:hi], minlength=B))
        p = perm.cpu().numpy().view(np.uint32)
        assert p.min() >= lo and p.max() < hi and np.unique(p).size == hi - lo     # global indices of this shard
        assert np.all(np.diff(keys[p].astype(np.int64)) >= 0)                      # grouped by bucket

        f = capi.unit_f32(n)
        idx = capi.fmix32(n, xor=0x85EBCA6B, mask=1023)
        bins = sh.scatter_add(torch.zeros(1024, device=dev), torch.from_numpy(f[lo:hi].copy()).to(dev),
                              torch.from_numpy(idx[lo:hi].view(np.int32).copy()).to(dev))
        exp = capi.scatter_reduce("f32", "add", np.zeros(1024, np.float32), f, idx, acc64=True)
        assert np.allclose(bins.cpu().numpy(), exp, rtol=1e-5)
        ft = torch.from_numpy(f[lo:hi].copy()).to(dev)
        d = sh.dot(ft, ft)
        ref = float(np.dot(f.astype(np.float64), f))
        assert abs(float(d[0]) - ref) <= 1e-6 * 22 * ref
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [(1 << 22) + 4096])
def test_sharded_primitives_nccl_world2(n):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), n), nprocs=2, join=True)
