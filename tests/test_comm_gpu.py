"""The sharded primitives with the combine step fused into their kernels (peer-memory communicator,
csrc/comm.cuh), all ranks inside ONE process: one thread and one device per rank, windows mapped
through plain peer access (drjit_b200_comm_connect_local -- the way Dr.Jit itself drives several
devices from one process, src/cuda_core.cpp:518-536).

world = 1 runs on any GPU box and still goes through every exchange code path (publish to the own
window, flag, wait, fold); world = 2 needs two devices with peer access and is skipped otherwise."""
import threading

import numpy as np
import pytest
import torch

from tests.dist_body import check_rank

pytestmark = pytest.mark.gpu


def _run_world(world, n, body):
    from drjit_b200.dist import PeerComm, Sharded
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    comms = PeerComm.local(list(range(world)), bulk_bytes=8 << 20)
    results, errors = [None] * world, []

    def worker(rank):
        try:
            torch.cuda.set_device(rank)
            dev = torch.device("cuda", rank)
            sh = Sharded(rank=rank, world=world, comm=comms[rank])
            results[rank] = body(sh, rank, world, dev, n)
            torch.cuda.synchronize(dev)
        except BaseException as e:  # noqa: BLE001
            errors.append((rank, e))

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    alive = [t for t in threads if t.is_alive()]
    assert not alive, "a rank is still blocked (exchange never completed)"
    if errors:
        raise errors[0][1]
    for c in comms:
        c.destroy()
    return results


@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("n", [(1 << 22) + 4096, 100_003])
def test_fused_sharded_primitives(world, n):
    bins = _run_world(world, n, check_rank)
    for b in bins[1:]:          # every bin folded once, in rank order: bit-identical on all ranks
        assert np.array_equal(b, bins[0])


@pytest.mark.parametrize("world", [1, 2])
def test_fused_empty_trailing_shard(world):
    """n < world: an empty shard still takes part in every exchange and contributes the identity"""
    from drjit_b200.ops import ReduceOp, VarType

    def body(sh, rank, world, dev, n):
        lo, hi = sh.shard_range(n)
        u = np.array([41], np.uint32)
        ut = torch.from_numpy(u[lo:hi].view(np.int32).copy()).to(dev)
        for op in (ReduceOp.Add, ReduceOp.Min, ReduceOp.Max, ReduceOp.Mul):
            assert sh.reduce(op, ut, vt=VarType.UInt32).cpu().numpy().view(np.uint32)[0] == 41
        got = sh.prefix_sum(ut, vt=VarType.UInt32).cpu().numpy().view(np.uint32)
        assert got.size == hi - lo and (got.size == 0 or got[0] == 0)
        _, off = sh.prefix_reduce_offsets(ReduceOp.Add, ut, vt=VarType.UInt32)
        assert int(off.cpu().numpy().view(np.uint32)[0]) == (0 if rank == 0 else 41)
        assert sh.all(torch.ones(hi - lo, dtype=torch.uint8, device=dev)) is True
        assert sh.any(torch.zeros(hi - lo, dtype=torch.uint8, device=dev)) is False
        _, counts = sh.compress(torch.ones(hi - lo, dtype=torch.uint8, device=dev), lo)
        assert counts == [1] + [0] * (world - 1)
        res = sh.mkperm(torch.zeros(hi - lo, dtype=torch.int32, device=dev), 4, lo)
        assert res.table.tolist() == [[0, 0, 1, 0]] and int(res.hist.sum()) == hi - lo
        f = torch.full((hi - lo,), 2.0, device=dev)
        assert float(sh.dot(f, f).cpu()[0]) == 4.0
        bins = sh.scatter_add(torch.zeros(8, device=dev), f, torch.zeros(hi - lo, dtype=torch.int32, device=dev))
        assert bins.cpu().tolist() == [2.0] + [0.0] * 7

    _run_world(world, 1, body)


@pytest.mark.parametrize("world", [1, 2])
def test_allreduce_bins_sizes_and_types(world):
    """drjit_b200_comm_allreduce: ragged sizes (scalar tail, empty slices), 4- and 8-byte types, back to
    back epochs; integer sums exact, results identical on every rank."""
    import ctypes
    from drjit_b200._lib import check, lib

    def body(sh, rank, world, dev, n):
        out = []
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for count in (1, 3, 4, 5, 1000, 4099, 1 << 20):
            for dt, vt in ((torch.int32, 8), (torch.float32, 14), (torch.int64, 10), (torch.float64, 15)):
                if count * dt.itemsize + 256 * world > (8 << 20):
                    continue
                g = torch.Generator(device="cpu").manual_seed(1234 + count)
                parts = [torch.randint(-1000, 1000, (count,), generator=g, dtype=torch.int64) for _ in range(world)]
                mine = parts[rank].to(dt).to(dev)
                check(lib.drjit_b200_comm_allreduce(sh.comm.ptr, stream, vt, 1, ctypes.c_void_p(mine.data_ptr()), count))
                exp = sum(parts).to(dt)         # small integers: exact in every type
                assert torch.equal(mine.cpu(), exp), (count, dt)
                out.append(mine.cpu())
        return out

    res = _run_world(world, 0, body)
    for r in res[1:]:
        assert all(torch.equal(a, b) for a, b in zip(r, res[0]))


@pytest.mark.parametrize("world", [1, 2])
def test_consecutive_one_sided_folds_never_overrun_a_slow_peer(world):
    """A lower-ranks fold needs no value on rank 0 and an upper-ranks fold none on the last rank. Every
    rank must still observe all peers in every epoch: with two cell parities a rank that ran two
    epochs ahead would overwrite a cell its peer has not read yet (the peer then waits for an epoch it
    can never see and traps after 20 s). Rank 0 enqueues 64 folds back to back while the other ranks
    are held back by a long-running kernel in front of theirs."""
    from drjit_b200.ops import ReduceOp, VarType

    def body(sh, rank, world, dev, n):
        K = 64
        src = [torch.tensor([(rank + 1) * (k + 1)], dtype=torch.int32, device=dev) for k in range(K)]
        low = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(K)]
        high = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(K)]
        torch.cuda.synchronize(dev)
        for lower, dst in ((True, low), (False, high)):
            slow = 0 if not lower else world - 1        # the rank whose values everybody else needs last
            if rank == slow and world > 1:
                torch.cuda._sleep(200_000_000)           # ~0.1 s on the stream in front of the folds
            for k in range(K):
                if lower:
                    sh.fold_scalar(ReduceOp.Add, src[k], dst[k], lower=True, vt=VarType.UInt32)
                else:   # upper-ranks fold: what a reverse scan uses (FOLD_HIGHER through the scan entry point)
                    x = torch.full((8,), (rank + 1) * (k + 1), dtype=torch.int32, device=dev)
                    import ctypes
                    from drjit_b200._lib import check, lib
                    out = torch.empty_like(x)
                    check(lib.drjit_b200_comm_prefix_reduce(
                        sh.comm.ptr, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream), int(VarType.UInt32),
                        int(ReduceOp.Add), 8, 1, 1, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                        ctypes.c_void_p(dst[k].data_ptr()), 0))
            torch.cuda.synchronize(dev)
        for k in range(K):
            assert int(low[k].cpu()[0]) == sum((r + 1) * (k + 1) for r in range(rank))
            assert int(high[k].cpu()[0]) == sum(8 * (r + 1) * (k + 1) for r in range(rank + 1, world))

    _run_world(world, 0, body)
