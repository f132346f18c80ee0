#!/usr/bin/env python
"""compute-sanitizer pass over the kernels added in round 2 / r5: group-per-block segmented scans
(G = 4 .. 32, forward and reverse, ragged ends, in place), the compaction's PEER instantiation (world 1:
the exchange runs against the own window), the flattened bins all-reduce entry point, call_reduce with
the payload tile staged in shared memory.

    compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_r5.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drjit_b200 as dr  # noqa: E402
from drjit_b200 import ReduceOp, VarType, ops  # noqa: E402
from drjit_b200.dist import PeerComm, Sharded  # noqa: E402

dev = torch.device("cuda", 0)
n = 300_017
for dt, vt in ((torch.int32, VarType.UInt32), (torch.float64, VarType.Float64), (torch.uint8, VarType.UInt8), (torch.float16, VarType.Float16)):
    x = (torch.rand(n, device=dev) * 50).to(dt)
    for bs in (12, 28, 64, 100, 1000, 4099):
        for ex, rev in ((True, False), (False, True)):
            out = ops.block_prefix_reduce(ReduceOp.Add, x, bs, ex, rev, vt=vt)
    y = x.clone()
    ops.block_prefix_reduce(ReduceOp.Max, y, 100, True, False, vt=vt, out=y)
torch.cuda.synchronize()

comm = PeerComm.local([0], bulk_bytes=8 << 20)[0]
sh = Sharded(rank=0, world=1, comm=comm)
m = torch.empty(1_000_003, dtype=torch.uint8, device=dev); ops.fill_fmix32(m, 2, and_=128)
out, counts = sh.compress(m, 0)
torch.cuda.synchronize()
assert counts[0] == int(m.sum().item())
out, counts = sh.compress(m[:0], 0)
assert counts == [0]
bins = sh.scatter_add(torch.zeros(4099, device=dev), torch.ones(100_000, device=dev),
                      torch.randint(0, 4099, (100_000,), dtype=torch.int32, device=dev))
torch.cuda.synchronize()
comm.destroy()

n = 148 * 2 * 1024 * 24 + 777
keys = torch.empty(n, dtype=torch.int32, device=dev); ops.fill_fmix32(keys, 0, and_=4095)
pay = torch.arange(n, dtype=torch.int32, device=dev)
perm, table, (po,) = dr.call_reduce(keys, 4096, [pay])
torch.cuda.synchronize()
assert torch.equal(po, perm.view(torch.int32))          # the payload was the index itself
print("sanitize_r5: done")
