#!/usr/bin/env python
"""compute-sanitizer pass over the code paths added after profiles/r1_sanitizer.txt: the unordered
mkperm tile kernels with 48 Ki / 40 Ki-key tiles (ragged last tile) and the f16 scatter-reductions.

    compute-sanitizer --tool memcheck python scripts/sanitize_new_paths.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drjit_b200 as dr  # noqa: E402
from drjit_b200 import ReduceMode, ReduceOp, ops  # noqa: E402

n = 148 * 2 * 48 * 1024 + 12_345
keys = torch.empty(n, dtype=torch.int32, device="cuda")
for buckets in (4096, 8192):
    ops.fill_fmix32(keys, 0, and_=buckets - 1)
    perm, table = dr.block_mkperm(keys, n, buckets)
    torch.cuda.synchronize()
    assert int(perm.view(torch.int32).to(torch.int64).sum().item()) == n * (n - 1) // 2, "not a permutation"

m = 100_003
idx = torch.empty(m, dtype=torch.int32, device="cuda"); ops.fill_fmix32(idx, 0, xor=3, and_=1023)
val = torch.empty(m, dtype=torch.float32, device="cuda"); ops.fill_fmix32(val, 1)
val16 = val.to(torch.float16)
for op in (ReduceOp.Add, ReduceOp.Min, ReduceOp.Max):
    for mode in (ReduceMode.Direct, ReduceMode.Local):
        tgt = torch.zeros(1025, dtype=torch.float16, device="cuda")
        dr.scatter_reduce(op, tgt[1:], val16, idx, mode=mode)      # odd start: partner halves on both sides
torch.cuda.synchronize()
print("sanitize_new_paths: done")
