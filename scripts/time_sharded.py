#!/usr/bin/env python
"""Per-primitive cost of the sharded forms in isolation (developer tool, run under torchrun):
R back-to-back calls of ONE primitive on the strong-scaling shard of the BASELINE size, device-timed,
max over ranks -- fused (peer-memory exchange inside the kernel) next to the NCCL path and to the
shard-local kernel alone. Separates what the exchange costs from host gaps and rank skew in bench.py.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/time_sharded.py [--reps R]
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drjit_b200 import ReduceOp, VarType, ops  # noqa: E402
from drjit_b200 import dist as ddist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    a = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = ddist.PeerComm.from_process_group(dist.group.WORLD, device=dev)
    fused = ddist.Sharded(rank, world, dist.group.WORLD, comm=comm)
    nccl = ddist.Sharded(rank, world, dist.group.WORLD)

    def timeit(fn):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.reps):
            fn()
        e.record(); torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / a.reps * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    n28, n30, n26 = (1 << 28) // world, (1 << 30) // world, (1 << 26) // world
    x = ops.fill_fmix32(torch.empty(n28, dtype=torch.float32, device=dev), 1, start=rank * n28)
    y = ops.fill_fmix32(torch.empty(n28, dtype=torch.float32, device=dev), 1, start=rank * n28, xor=5)
    u = ops.fill_fmix32(torch.empty(n30, dtype=torch.int32, device=dev), 0, start=rank * n30)
    uo = torch.empty_like(u)
    m = ops.fill_fmix32(torch.empty(n30, dtype=torch.uint8, device=dev), 2, start=rank * n30, and_=128)
    co = torch.empty(n30, dtype=torch.int32, device=dev)
    k = ops.fill_fmix32(torch.empty(n26, dtype=torch.int32, device=dev), 0, start=rank * n26, and_=4095)
    perm = torch.empty_like(k)
    si = ops.fill_fmix32(torch.empty(n28, dtype=torch.int32, device=dev), 0, start=rank * n28, xor=0x85EBCA6B, and_=(1 << 20) - 1)
    bins = torch.zeros(1 << 20, dtype=torch.float32, device=dev)
    one = torch.ones(1, dtype=torch.int32, device=dev); tmp = torch.zeros(1, dtype=torch.int32, device=dev)
    o1 = torch.empty(1, dtype=torch.float32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    hist = torch.empty(4096, dtype=torch.int32, device=dev); rb = torch.empty(4096, dtype=torch.int32, device=dev)

    rows = []

    def row(name, local_fn, fused_fn, nccl_fn):
        rows.append((name, timeit(local_fn) if local_fn else float("nan"), timeit(fused_fn), timeit(nccl_fn)))

    row("exchange only (fold of one scalar)", None, lambda: fused.fold_scalar(ReduceOp.Add, one, tmp), lambda: nccl.fold_scalar(ReduceOp.Add, one, tmp))
    row("sum f32", lambda: ops.block_reduce(ReduceOp.Add, x, n28, out=o1), lambda: fused.reduce(ReduceOp.Add, x, out=o1), lambda: nccl.reduce(ReduceOp.Add, x))
    row("dot f32", lambda: ops.dot(x, y), lambda: fused.dot(x, y, out=o1), lambda: nccl.dot(x, y))
    row("prefix_sum u32 (offset form)", lambda: ops.block_prefix_reduce(ReduceOp.Add, u, n30, True, False, vt=VarType.UInt32, out=uo),
        lambda: fused.prefix_reduce_offsets(ReduceOp.Add, u, vt=VarType.UInt32, out=uo, offset=tmp),
        lambda: nccl.prefix_reduce_offsets(ReduceOp.Add, u, vt=VarType.UInt32, out=uo))
    row("compress (synchronous)", lambda: ops.compress_async(m, 0, out=co, count=cnt), lambda: fused.compress(m, 0, out=co), lambda: nccl.compress(m, 0, out=co))
    row("mkperm 4096 (+ global table)", lambda: ops.mkperm_sharded(k, 4096, 0, perm=perm, hist=hist),
        lambda: fused.mkperm(k, 4096, 0, perm=perm, hist=hist, rank_base=rb, raw_table=True), lambda: nccl.mkperm(k, 4096, 0, perm=perm))
    row("scatter_add -> 2^20 bins", lambda: ops.scatter_reduce(ReduceOp.Add, bins, x, si), lambda: fused.scatter_add(bins, x, si), lambda: nccl.scatter_add(bins, x, si))
    row("all-reduce of the 4 MB bins alone", None, lambda: fused.scatter_add(bins, x[:0], si[:0]), lambda: nccl.scatter_add(bins, x[:0], si[:0]))
    if rank == 0:
        print(f"world {world}: us per call, {a.reps} back-to-back calls, max over ranks")
        print(f"{'primitive':40s} {'shard-local':>12s} {'fused':>10s} {'nccl':>10s}")
        for name, l, f, n in rows:
            print(f"{name:40s} {l:12.1f} {f:10.1f} {n:10.1f}")
    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
