#!/usr/bin/env python
"""Throughput of the less-travelled variants (block sizes, element types, ops) at 2^28 elements (2^27
for 8-byte types): a quick way to spot slow paths. GB/s are algorithmic bytes (read + written)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drjit_b200 import ReduceOp, VarType, ops  # noqa: E402


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = "cuda"
    n = 1 << 28
    f32 = torch.empty(n, dtype=torch.float32, device=dev); ops.fill_fmix32(f32, 1)
    print("block_reduce(Add) f32, 2^28 elements")
    for bs in (2, 3, 4, 7, 16, 100, 256, 1000, 4096, 65536, 1 << 20, n):
        out = torch.empty((n + bs - 1) // bs, dtype=torch.float32, device=dev)
        ms = timeit(lambda: ops.block_reduce(ReduceOp.Add, f32, bs, out=out))
        print(f"  block_size {bs:>10d}: {ms:7.3f} ms  {(n * 4 + out.numel() * 4) / ms / 1e6:7.1f} GB/s")
    print("full reductions, other types / ops")
    cases = [("u8 add", torch.uint8, VarType.UInt8, ReduceOp.Add, n), ("f16 add", torch.float16, VarType.Float16, ReduceOp.Add, n),
             ("i32 min", torch.int32, VarType.Int32, ReduceOp.Min, n), ("u32 or", torch.int32, VarType.UInt32, ReduceOp.Or, n),
             ("f32 max", torch.float32, VarType.Float32, ReduceOp.Max, n), ("f64 add", torch.float64, VarType.Float64, ReduceOp.Add, n // 2),
             ("u64 add", torch.int64, VarType.UInt64, ReduceOp.Add, n // 2)]
    for name, dt, vt, op, m in cases:
        x = torch.ones(m, dtype=dt, device=dev)
        ms = timeit(lambda: ops.block_reduce(op, x, m, vt=vt))
        print(f"  {name:8s}: {ms:7.3f} ms  {m * x.element_size() / ms / 1e6:7.1f} GB/s")
        del x
    print("prefix reductions (inclusive), other types / block sizes")
    cases = [("u8 add", torch.uint8, VarType.UInt8, ReduceOp.Add, n, n), ("f16 add", torch.float16, VarType.Float16, ReduceOp.Add, n, n),
             ("f32 add", torch.float32, VarType.Float32, ReduceOp.Add, n, n), ("f32 max", torch.float32, VarType.Float32, ReduceOp.Max, n, n),
             ("f64 add", torch.float64, VarType.Float64, ReduceOp.Add, n // 2, n // 2),
             ("f32 add bs=4", torch.float32, VarType.Float32, ReduceOp.Add, n, 4), ("f32 add bs=256", torch.float32, VarType.Float32, ReduceOp.Add, n, 256),
             ("f32 add bs=2^20", torch.float32, VarType.Float32, ReduceOp.Add, n, 1 << 20), ("u32 add reverse", torch.int32, VarType.UInt32, ReduceOp.Add, n, n)]
    for name, dt, vt, op, m, bs in cases:
        x = torch.ones(m, dtype=dt, device=dev); out = torch.empty_like(x)
        rev = "reverse" in name
        ms = timeit(lambda: ops.block_prefix_reduce(op, x, bs, False, rev, vt=vt, out=out))
        print(f"  {name:16s}: {ms:7.3f} ms  {2 * m * x.element_size() / ms / 1e6:7.1f} GB/s")
        del x, out
    print("dr.all / dr.any over 2^30 mask bytes")
    mask = torch.ones(1 << 30, dtype=torch.uint8, device=dev)
    from drjit_b200 import ops as o
    for name, fn in (("all", o.all), ("any", o.any)):
        ms = timeit(lambda: fn(mask))
        print(f"  {name}: {ms:7.3f} ms  {(1 << 30) / ms / 1e6:7.1f} GB/s")


if __name__ == "__main__":
    main()
