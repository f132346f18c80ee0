#!/usr/bin/env python
"""Segmented prefix sums (dr.block_prefix_sum) over block sizes and types: time per call and
fraction of the measured copy bandwidth (read + write of the array)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drjit_b200 as dr          # noqa: E402
from drjit_b200 import ReduceOp, VarType, ops   # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = "cuda"
    print(f"{'type':5s} {'block':>8s} {'ms':>8s} {'GB/s':>8s} {'of peak':>8s}")
    for name, dt, vt, n in (("f32", torch.float32, VarType.Float32, 1 << 28), ("u32", torch.int32, VarType.UInt32, 1 << 28),
                            ("f64", torch.float64, VarType.Float64, 1 << 27), ("u8", torch.uint8, VarType.UInt8, 1 << 30),
                            ("f16", torch.float16, VarType.Float16, 1 << 29)):
        x = torch.rand(n, dtype=dt, device=dev) if dt.is_floating_point else torch.randint(0, 100, (n,), dtype=dt, device=dev)
        out = torch.empty_like(x)
        for bs in (9, 12, 16, 20, 24, 28, 33, 40, 48, 64, 80, 100, 128, 160, 256, 1000, 4096, 32768 // x.element_size() * 4, 100000, 1 << 20):
            for ex, rev in ((True, False),):
                ms = timeit(lambda: ops.block_prefix_reduce(ReduceOp.Add, x, bs, ex, rev, vt=vt, out=out))
                gbs = 2 * n * x.element_size() / ms / 1e6
                print(f"{name:5s} {bs:8d} {ms:8.3f} {gbs:8.1f} {100 * gbs / PEAK:7.1f}%  {'excl' if ex else 'incl'}{' rev' if rev else ''}", flush=True)
        del x, out


if __name__ == "__main__":
    main()
