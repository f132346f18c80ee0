"""One-process-per-GPU sharding of the primitive path.

The reference is single-device (no collectives anywhere in the tree). Here every primitive
runs shard-locally on a contiguous index range [lo, hi) of the global array and the shards
are combined with ONE tiny collective each -- an all-gather of per-rank scalars / counts, an
all-reduce of a 16 KB histogram or of the 4 MB bin array. torch.distributed (NCCL on GPUs,
gloo in the CPU tests) is only the plumbing; the combine arithmetic itself (the scan of the
gathered totals, the final reduction of the gathered partials) reuses the same kernels.

``Sharded(local=...)`` takes the object that supplies the shard-local primitives. The default
is ``drjit_b200.ops`` (CUDA, no fallback); the CPU tests inject an oracle-backed stand-in to
exercise the host-side logic under gloo with world_size 2.
"""
import torch
import torch.distributed as dist

from .ops import ReduceOp


def shard_bounds(n, world, align=1):
    """Contiguous, `align`-aligned shard boundaries: list of world+1 offsets covering [0, n)."""
    per = -(-n // world)                       # ceil
    per = -(-per // align) * align
    return [min(r * per, n) for r in range(world + 1)]


class Sharded:
    def __init__(self, rank=0, world=1, group=None, local=None):
        if local is None:
            from . import ops as local          # CUDA path; import fails loudly without the library
        self.rank, self.world, self.group, self.local = rank, world, group, local

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

    # ------------------------------------------------------------------ reductions
    def reduce(self, op, x, vt=None):
        """dr.sum/prod/min/max over the global array: local reduce -> all-gather of W partials ->
        the same kernel folds them (type- and op-exact for unsigned types, which NCCL lacks)."""
        part = self.local.block_reduce(op, x, x.numel(), vt=vt)
        if self.world == 1:
            return part
        allp = self._all_gather(part)
        return self.local.block_reduce(op, allp, allp.numel(), vt=vt)

    def block_reduce(self, op, x, block_size, vt=None, out=None):
        """dr.block_reduce over the global array. Shards cut at multiples of ``block_size``
        (``shard_range(n, align=block_size)``) hold whole blocks, so every rank reduces its own
        blocks and the global result is the rank-order concatenation: no collective."""
        return self.local.block_reduce(op, x, block_size, vt=vt, out=out)

    def _all_any(self, mask, want_all):
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
        """dr.all over the global mask (synchronous like jitc_all): local flag, one 4-byte all-reduce (min)."""
        return self._all_any(mask, True)

    def any(self, mask):    # noqa: A003
        """dr.any over the global mask: local flag, one 4-byte all-reduce (max)."""
        return self._all_any(mask, False)

    def dot(self, a, b):
        part = self.local.dot(a, b)
        if self.world == 1:
            return part
        allp = self._all_gather(part)
        return self.local.block_reduce(ReduceOp.Add, allp, allp.numel())

    # ------------------------------------------------------------------ prefix sum
    def prefix_reduce(self, op, x, exclusive=True, vt=None, out=None):
        """Global prefix reduction. Rank r needs op-reduction of all lower shards as its carry:
        local total (one read pass) -> all-gather -> exclusive scan of the W totals (same
        kernel) -> shard scan seeded with carry[r]. Data is read twice and written once; the
        only inter-GPU traffic is W scalars."""
        if self.world == 1:
            return self.local.block_prefix_reduce(op, x, x.numel(), exclusive, False, vt=vt, out=out)
        total = self.local.block_reduce(op, x, x.numel(), vt=vt)
        totals = self._all_gather(total)
        carries = self.local.block_prefix_reduce(op, totals, totals.numel(), True, False, vt=vt)
        return self.local.prefix_reduce_carry(op, x, exclusive, False, carry_in=carries[self.rank:self.rank + 1],
                                              vt=vt, out=out)

    def prefix_sum(self, x, exclusive=True, vt=None, out=None):
        return self.prefix_reduce(ReduceOp.Add, x, exclusive, vt, out)

    def prefix_reduce_offsets(self, op, x, exclusive=True, vt=None, out=None):
        """Global prefix reduction in *shard-offset form*: returns ``(local, offset)`` where ``local``
        is the prefix reduction of this rank's shard alone and ``offset`` (1-element device tensor)
        the reduction of all lower shards, i.e. global[i] = op(offset, local[i]). This is the form
        the compress / mkperm offsets take as well: one pass over the shard (read + write, the
        shard total falls out of the same kernel), then an all-gather of W totals and a W-element
        exclusive scan. A consumer folds ``offset`` into whatever it does with the values; the
        materialised form (``prefix_reduce``) costs one more read pass over the shard."""
        total = torch.empty(1, dtype=x.dtype, device=x.device)
        local = self.local.prefix_reduce_carry(op, x, exclusive, False, carry_in=None, total_out=total,
                                               vt=vt, out=out)
        if self.world == 1:
            ident = self.local.block_prefix_reduce(op, total, 1, True, False, vt=vt)   # exclusive scan of 1 = identity
            return local, ident
        totals = self._all_gather(total)
        carries = self.local.block_prefix_reduce(op, totals, totals.numel(), True, False, vt=vt)
        return local, carries[self.rank:self.rank + 1]

    # ------------------------------------------------------------------ compress
    def compress(self, mask, index_base, out=None):
        """Shard-local compaction with global indices. Returns (out, counts): rank r owns
        out[:counts[r]]; the global list is the rank-order concatenation. Like the reference
        (cuda_ts.cpp:759) the call ends with one host synchronisation to read the counts."""
        out, count = self.local.compress_async(mask, index_base, out=out)
        counts = self._all_gather(count)
        return out, [int(c) & 0xFFFFFFFF for c in counts.cpu().tolist()]

    # ------------------------------------------------------------------ mkperm
    def mkperm(self, keys, bucket_count, index_base, perm=None):
        """Shard-local permutation (entries are global indices) + global bucket histogram.
        Returns (perm, local_hist, global_hist_host): bucket b of the global, rank-major
        stable order is the concatenation over ranks of perm[start_r[b] : start_r[b] + local_hist[b]]."""
        perm, hist = self.local.mkperm_sharded(keys, bucket_count, index_base, perm=perm)
        ghist = hist
        if self.world > 1:
            ghist = hist.clone()
            dist.all_reduce(ghist, op=dist.ReduceOp.SUM, group=self.group)
        # the vcall dispatcher needs the table of non-empty buckets on the host (call.cpp:1346-1378)
        return perm, hist, ghist.cpu()

    # ------------------------------------------------------------------ scatter-add
    def scatter_add(self, bins, value, index):
        """Each rank accumulates its shard into its own bin array; bins are all-reduced."""
        self.local.scatter_reduce(ReduceOp.Add, bins, value, index)
        if self.world > 1:
            dist.all_reduce(bins, op=dist.ReduceOp.SUM, group=self.group)
        return bins
