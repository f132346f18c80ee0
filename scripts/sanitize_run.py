#!/usr/bin/env python
"""One pass over every primitive at sizes that exercise the multi-tile paths, for use under
compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python scripts/sanitize_run.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drjit_b200 as dr  # noqa: E402
from drjit_b200 import ReduceOp, VarType, ops  # noqa: E402

n = (1 << 22) + 4321            # > 148 * 2 tiles of 32 KiB: several iterations per CTA for u8 / masks
x = torch.empty(n, dtype=torch.int32, device="cuda"); ops.fill_fmix32(x, 0)
f = torch.empty(n, dtype=torch.float32, device="cuda"); ops.fill_fmix32(f, 1)
big = torch.empty(1 << 24, dtype=torch.int32, device="cuda"); ops.fill_fmix32(big, 0)   # 2048 scan tiles
m = torch.empty(1 << 25, dtype=torch.uint8, device="cuda"); ops.fill_fmix32(m, 2, and_=128)  # 1024 mask tiles
dr.sum(f); dr.dot(f, f); ops.block_reduce(ReduceOp.Add, f, 256)
ops.block_prefix_reduce(ReduceOp.Add, big, big.numel(), True, False, vt=VarType.UInt32)
ops.block_prefix_reduce(ReduceOp.Add, x, n, False, True, vt=VarType.UInt32)
ops.block_prefix_reduce(ReduceOp.Add, f, 1000, True, False)
ops.block_prefix_reduce(ReduceOp.Max, big.to(torch.int64), big.numel(), True, False, vt=VarType.Int64)
dr.compress(m); dr.compress(m[3:n])
for buckets in (64, 1000, 4096):
    keys = (x.to(torch.int64) & 0xFFFFFFFF).remainder(buckets).to(torch.int32)
    dr.block_mkperm(keys, n, buckets)
    dr.block_mkperm(keys[:50000], 50000, buckets)
    dr.block_mkperm(keys[:50000], 1000, buckets)
idx = (x.to(torch.int64) & 0xFFFFF).to(torch.int32)
dr.scatter_add(torch.zeros(1 << 20, device="cuda"), f, idx)
torch.cuda.synchronize()
print("sanitize_run: done")
