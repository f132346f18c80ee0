"""One-process-per-GPU sharding of the primitive path.

The reference is single-device (no collectives anywhere in the tree). Here every primitive runs on
a contiguous shard [lo, hi) of the global array and the shards are combined with ONE small exchange
each: a scalar, W counts, a bucket histogram, the 4 MB bin array.

Two ways to run that exchange:

* ``Sharded(comm=PeerComm...)`` -- the product path on GPUs. The exchange is fused into the
  primitive's own kernel: its last CTA stores the partial into every peer's window over NVLink,
  raises a flag and folds the W partials in rank order (``csrc/comm.cuh``). No library collective,
  no second launch, no host round trip; torch.distributed is used once, to pass the 64-byte window
  handles around when the communicator is created.
* ``Sharded(group=...)`` without a communicator -- the same host logic over torch.distributed
  collectives (NCCL on GPUs, gloo in the CPU tests, where an oracle-backed stand-in supplies the
  shard-local primitives). This is the library-collective baseline the fused path is measured
  against (bench.py reports both) and what the CPU test-suite exercises with world_size 2.
"""
import collections
import ctypes

import torch
import torch.distributed as dist

from .ops import ReduceOp, VarType, _on

FOLD_ALL, FOLD_LOWER, FOLD_HIGHER = 0, 1, 2
COMM_HANDLE_BYTES = 64
COMM_MAX_BUCKETS = 16384

MkpermResult = collections.namedtuple("MkpermResult", "perm hist rank_base table")
MkpermResult.__doc__ = """perm: shard permutation (device, entries are global indices); hist: shard counts per bucket
(device); rank_base: start of this rank's keys of every bucket in the global rank-major (= stable)
order (device); table: host int64 tensor (unique, 4) with rows {bucket id, global start, global size, 0}."""


def shard_bounds(n, world, align=1):
    """Contiguous, `align`-aligned shard boundaries: list of world+1 offsets covering [0, n)."""
    per = -(-n // world)                       # ceil
    per = -(-per // align) * align
    return [min(r * per, n) for r in range(world + 1)]


def _identity(op, like, vt=None):
    """1-element tensor holding the identity of `op` for arrays like `like` (src/var.cpp:2642-2652)."""
    dt = like.dtype
    unsigned = vt in (VarType.UInt32, VarType.UInt64, VarType.UInt8) or dt in (torch.uint8, torch.bool)
    if op in (ReduceOp.Add, ReduceOp.Or):
        v = 0
    elif op == ReduceOp.Mul:
        v = 1
    elif op == ReduceOp.And:
        v = -1 if dt not in (torch.uint8, torch.bool) else 255
    elif dt.is_floating_point:
        v = float("inf") if op == ReduceOp.Min else float("-inf")
    elif op == ReduceOp.Min:
        v = (255 if dt == torch.uint8 else -1) if unsigned else torch.iinfo(dt).max
    else:
        v = 0 if unsigned else torch.iinfo(dt).min
    return torch.full((1,), v, dtype=dt, device=like.device)


class PeerComm:
    """Peer-memory communicator of one rank (C ABI ``drjit_b200_comm_*``, ``csrc/comm.cu``)."""

    def __init__(self, rank, world, device=None, bulk_bytes=8 << 20):
        from ._lib import check, lib
        self._lib, self._check = lib, check
        self.rank, self.world, self.bulk_bytes = rank, world, bulk_bytes
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        handle = ctypes.c_void_p()
        with _on(self.device):
            check(lib.drjit_b200_comm_create(rank, world, bulk_bytes, ctypes.byref(handle)))
        self.ptr = handle

    def handle_bytes(self):
        buf = ctypes.create_string_buffer(COMM_HANDLE_BYTES)
        with _on(self.device):
            self._check(self._lib.drjit_b200_comm_handle(self.ptr, buf))
        return buf.raw

    def connect(self, handles):
        """handles: the `world` window handles in rank order (bytes)"""
        blob = b"".join(handles)
        assert len(blob) == self.world * COMM_HANDLE_BYTES
        with _on(self.device):
            self._check(self._lib.drjit_b200_comm_connect(self.ptr, ctypes.create_string_buffer(blob, len(blob))))
        return self

    @classmethod
    def from_process_group(cls, group=None, device=None, bulk_bytes=8 << 20):
        """One process per GPU: window handles travel once through torch.distributed."""
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        comm = cls(rank, world, device, bulk_bytes)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, comm.handle_bytes(), group=group)
            comm.connect(handles)
            dist.barrier(group=group)
        return comm

    @classmethod
    def local(cls, devices, bulk_bytes=8 << 20):
        """All ranks inside this process (rank r on devices[r]); returns the list of communicators."""
        from ._lib import check, lib
        world = len(devices)
        comms = [cls(r, world, torch.device("cuda", d) if isinstance(d, int) else d, bulk_bytes)
                 for r, d in enumerate(devices)]
        arr = (ctypes.c_void_p * world)(*[c.ptr for c in comms])
        check(lib.drjit_b200_comm_connect_local(arr, world))
        return comms

    def destroy(self):
        if self.ptr:
            self._lib.drjit_b200_comm_destroy(self.ptr)
            self.ptr = None


class Sharded:
    def __init__(self, rank=0, world=1, group=None, local=None, comm=None):
        if local is None:
            from . import ops as local          # CUDA path; import fails loudly without the library
        self.rank, self.world, self.group, self.local, self.comm = rank, world, group, local, comm
        if comm is not None:
            from ._lib import check, lib
            self._lib, self._check = lib, check
            assert comm.rank == rank and comm.world == world

    # ------------------------------------------------------------------ helpers
    def shard_range(self, n, align=1):
        b = shard_bounds(n, self.world, align)
        return b[self.rank], b[self.rank + 1]

    def _all_gather(self, t):
        """[W * len(t)] tensor holding every rank's `t` in rank order (device-side, no host sync)."""
        if self.world == 1:
            return t
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    @staticmethod
    def _s(x):
        return ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)

    @staticmethod
    def _p(x):
        return ctypes.c_void_p(x.data_ptr()) if x is not None else None

    def _vt(self, x, vt):
        from .ops import _vt
        return int(_vt(x, vt))

    # ------------------------------------------------------------------ reductions
    def reduce(self, op, x, vt=None, out=None):
        """dr.sum/prod/min/max over the global array. Fused path: ONE launch (shard reduction whose
        last CTA exchanges the partials over NVLink and folds them in rank order). Collective
        path: local reduce -> all-gather of W partials -> the same kernel folds them (type- and
        op-exact for unsigned types, which NCCL lacks)."""
        if self.comm is not None:
            if out is None:
                out = torch.empty(1, dtype=x.dtype, device=x.device)
            with _on(x.device):
                self._check(self._lib.drjit_b200_comm_reduce(self.comm.ptr, self._s(x), self._vt(x, vt), int(op),
                                                             FOLD_ALL, x.numel(), self._p(x), self._p(out)))
            return out
        part = self.local.block_reduce(op, x, x.numel(), vt=vt) if x.numel() else _identity(op, x, vt)
        if self.world == 1:
            return part
        allp = self._all_gather(part)
        return self.local.block_reduce(op, allp, allp.numel(), vt=vt)

    def block_reduce(self, op, x, block_size, vt=None, out=None):
        """dr.block_reduce over the global array. Shards cut at multiples of ``block_size``
        (``shard_range(n, align=block_size)``) hold whole blocks, so every rank reduces its own
        blocks and the global result is the rank-order concatenation: no exchange."""
        return self.local.block_reduce(op, x, block_size, vt=vt, out=out)

    def _all_any(self, mask, want_all):
        if self.comm is not None:
            res = ctypes.c_int(0)
            fn = self._lib.drjit_b200_comm_all if want_all else self._lib.drjit_b200_comm_any
            with _on(mask.device):
                self._check(fn(self.comm.ptr, self._s(mask), self._p(mask), mask.numel(), ctypes.byref(res)))
            return bool(res.value)
        if mask.numel() == 0:
            flag = want_all                     # identity of And / Or (empty trailing shard)
        else:
            flag = self.local.all(mask) if want_all else self.local.any(mask)
        if self.world == 1:
            return bool(flag)
        t = torch.tensor([int(bool(flag))], dtype=torch.int32, device=mask.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN if want_all else dist.ReduceOp.MAX, group=self.group)
        return bool(int(t.item()))

    def all(self, mask):    # noqa: A003
        """dr.all over the global mask (synchronous like jitc_all)."""
        return self._all_any(mask, True)

    def any(self, mask):    # noqa: A003
        """dr.any over the global mask."""
        return self._all_any(mask, False)

    def dot(self, a, b, out=None):
        if self.comm is not None:
            if out is None:
                out = torch.empty(1, dtype=a.dtype, device=a.device)
            with _on(a.device):
                self._check(self._lib.drjit_b200_comm_reduce_dot(self.comm.ptr, self._s(a), self._vt(a, None),
                                                                 self._p(a), self._p(b), a.numel(), self._p(out)))
            return out
        part = self.local.dot(a, b) if a.numel() else torch.zeros(1, dtype=a.dtype, device=a.device)
        if self.world == 1:
            return part
        allp = self._all_gather(part)
        return self.local.block_reduce(ReduceOp.Add, allp, allp.numel())

    def fold_scalar(self, op, src, dst, lower=False, vt=None):
        """dst[0] = fold over the ranks (``lower``: only the ranks below the caller, the carry of a
        forward scan) of every rank's 1-element device tensor ``src``. Fused path: one tiny
        peer-exchange kernel; collective path: all-gather + local fold."""
        if self.comm is not None:
            with _on(src.device):
                self._check(self._lib.drjit_b200_comm_fold(self.comm.ptr, self._s(src), self._vt(src, vt), int(op),
                                                           FOLD_LOWER if lower else FOLD_ALL, self._p(src), self._p(dst)))
            return dst
        allp = self._all_gather(src)
        if lower:
            allp = allp[:self.rank] if self.rank else _identity(op, src, vt)
        dst.copy_(self.local.block_reduce(op, allp, allp.numel(), vt=vt))
        return dst

    # ------------------------------------------------------------------ prefix sum
    def prefix_reduce(self, op, x, exclusive=True, vt=None, out=None):
        """Global prefix reduction, *materialised*: every element carries the global value. Rank r needs
        the reduction of all lower shards as its carry before it can emit anything, so the shard is
        read twice (12 B/element for 4-byte types). Fused path: shard reduction with the totals
        exchanged inside its last CTA, then the scan seeded with the carry -- two launches."""
        if out is None:
            out = torch.empty_like(x)
        if self.comm is not None:
            with _on(x.device):
                self._check(self._lib.drjit_b200_comm_prefix_reduce(
                    self.comm.ptr, self._s(x), self._vt(x, vt), int(op), x.numel(), int(exclusive), 0,
                    self._p(x), self._p(out), None, 1))
            return out
        if self.world == 1:
            return self.local.block_prefix_reduce(op, x, x.numel(), exclusive, False, vt=vt, out=out)
        total = self.local.block_reduce(op, x, x.numel(), vt=vt) if x.numel() else _identity(op, x, vt)
        totals = self._all_gather(total)
        carries = self.local.block_prefix_reduce(op, totals, totals.numel(), True, False, vt=vt)
        if x.numel() == 0:
            return out
        return self.local.prefix_reduce_carry(op, x, exclusive, False, carry_in=carries[self.rank:self.rank + 1],
                                              vt=vt, out=out)

    def prefix_sum(self, x, exclusive=True, vt=None, out=None):
        return self.prefix_reduce(ReduceOp.Add, x, exclusive, vt, out)

    def prefix_reduce_offsets(self, op, x, exclusive=True, vt=None, out=None, offset=None):
        """Global prefix reduction in *shard-offset form*: returns ``(local, offset)`` where ``local``
        is the prefix reduction of this rank's shard alone and ``offset`` (1-element device tensor)
        the reduction of all lower shards, i.e. global[i] = op(offset, local[i]). This is the form
        the compress / mkperm offsets take as well: one pass over the shard (read + write, the
        shard total falls out of the same kernel), then the exchange of W totals. A consumer folds
        ``offset`` into whatever it does with the values."""
        if out is None:
            out = torch.empty_like(x)
        if self.comm is not None:
            if offset is None:
                offset = torch.empty(1, dtype=x.dtype, device=x.device)
            with _on(x.device):
                self._check(self._lib.drjit_b200_comm_prefix_reduce(
                    self.comm.ptr, self._s(x), self._vt(x, vt), int(op), x.numel(), int(exclusive), 0,
                    self._p(x), self._p(out), self._p(offset), 0))
            return out, offset
        total = _identity(op, x, vt)
        local = out
        if x.numel():
            local = self.local.prefix_reduce_carry(op, x, exclusive, False, carry_in=None, total_out=total,
                                                   vt=vt, out=out)
        if self.world == 1:
            return local, _identity(op, x, vt)
        totals = self._all_gather(total)
        carries = self.local.block_prefix_reduce(op, totals, totals.numel(), True, False, vt=vt)
        return local, carries[self.rank:self.rank + 1]

    # ------------------------------------------------------------------ compress
    def compress(self, mask, index_base, out=None):
        """Shard-local compaction with global indices. Returns (out, counts): rank r owns
        out[:counts[r]]; the global list is the rank-order concatenation. The call blocks until the
        counts are known. Fused path: ONE launch -- the thread of the compaction kernel that learns the
        shard's count exchanges it over NVLink and writes all W counts into pinned host memory, the
        host spins on a sequence word; like jit_block_mkperm (cuda_ts.cpp:953-967) the call returns
        once the host-visible part of the result is there, ``out`` is complete in stream order.
        Collective path: compress_async + all-gather + one host synchronisation."""
        if self.comm is not None:
            if out is None:
                out = torch.empty(mask.numel(), dtype=torch.int32, device=mask.device)
            counts = (ctypes.c_uint32 * self.world)()
            with _on(mask.device):
                self._check(self._lib.drjit_b200_comm_compress(self.comm.ptr, self._s(mask), self._p(mask),
                                                               mask.numel(), index_base & 0xFFFFFFFF, self._p(out), counts))
            return out, [int(c) for c in counts]
        if mask.numel():
            out, count = self.local.compress_async(mask, index_base, out=out)
        else:                                   # empty trailing shard
            out = torch.empty(0, dtype=torch.int32, device=mask.device) if out is None else out
            count = torch.zeros(1, dtype=torch.int32, device=mask.device)
        counts = self._all_gather(count)
        return out, [int(c) & 0xFFFFFFFF for c in counts.cpu().tolist()]

    # ------------------------------------------------------------------ mkperm
    def mkperm(self, keys, bucket_count, index_base, perm=None, want_table=True, hist=None, rank_base=None,
               raw_table=False):
        """Shard-local permutation (entries are global indices) + the GLOBAL bucket table.
        Bucket b of the global, rank-major (= stable) order is the concatenation over ranks of
        perm_r[local_start_r[b] : local_start_r[b] + hist_r[b]], and rank r's piece begins at
        rank_base_r[b] of the global permutation. Returns ``MkpermResult``. Fused path: the shard
        histograms are exchanged inside the bucket-scan kernel and the table lands in pinned host
        memory; collective path: all-gather of the histograms + host arithmetic."""
        dev = keys.device
        if self.comm is not None and bucket_count <= COMM_MAX_BUCKETS:
            from .ops import _pinned_offsets
            if perm is None:
                perm = torch.empty(keys.numel(), dtype=torch.int32, device=dev)
            if hist is None:
                hist = torch.empty(bucket_count, dtype=torch.int32, device=dev)
            if rank_base is None:
                rank_base = torch.empty(bucket_count, dtype=torch.int32, device=dev)
            offsets = _pinned_offsets(bucket_count) if want_table else None
            unique = ctypes.c_uint32(0)
            with _on(dev):
                self._check(self._lib.drjit_b200_comm_mkperm(
                    self.comm.ptr, self._s(keys), self._p(keys), keys.numel(), bucket_count, index_base & 0xFFFFFFFF,
                    self._p(perm), self._p(hist), self._p(rank_base), self._p(offsets), ctypes.byref(unique)))
            table = None
            if offsets is not None and raw_table:       # the pinned buffer as the call left it (reused by the next call)
                table = (offsets, unique.value)
            elif offsets is not None:
                table = offsets[:4 * unique.value].clone().view(-1, 4).to(torch.int64) & 0xFFFFFFFF
            return MkpermResult(perm, hist, rank_base, table)
        if keys.numel():
            perm, hist = self.local.mkperm_sharded(keys, bucket_count, index_base, perm=perm)
        else:                                   # empty trailing shard
            perm = torch.empty(0, dtype=torch.int32, device=dev) if perm is None else perm
            hist = torch.zeros(bucket_count, dtype=torch.int32, device=dev)
        allh = self._all_gather(hist).view(self.world, bucket_count).to(torch.int64) & 0xFFFFFFFF
        gsize = allh.sum(0)
        gstart = torch.cumsum(gsize, 0) - gsize
        rank_base = (gstart + allh[:self.rank].sum(0)).to(torch.int32)
        table = None
        if want_table:      # the vcall dispatcher needs the non-empty buckets on the host (call.cpp:1346-1378)
            ids = torch.nonzero(gsize).flatten()
            table = torch.stack([ids, gstart[ids], gsize[ids], torch.zeros_like(ids)], 1).cpu()
        return MkpermResult(perm, hist, rank_base, table)

    # ------------------------------------------------------------------ scatter-add
    def scatter_add(self, bins, value, index):
        """Each rank accumulates its shard into its own bin array; the bins are then summed over
        the ranks in place. Fused path: one peer-memory kernel (reduce-scatter + all-gather through
        the windows, every bin folded once in rank order); collective path: NCCL all-reduce."""
        if value.numel():
            self.local.scatter_reduce(ReduceOp.Add, bins, value, index)
        if self.world == 1:
            return bins
        if self.comm is not None and bins.numel() * bins.element_size() + 256 * self.world <= self.comm.bulk_bytes:
            with _on(bins.device):
                self._check(self._lib.drjit_b200_comm_allreduce(self.comm.ptr, self._s(bins), self._vt(bins, None),
                                                                int(ReduceOp.Add), self._p(bins), bins.numel()))
            return bins
        dist.all_reduce(bins, op=dist.ReduceOp.SUM, group=self.group)
        return bins
