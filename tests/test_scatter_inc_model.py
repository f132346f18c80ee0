"""CPU models of the control flow of the scatter_inc kernels (drjit_b200/csrc/scatter_packet.cu): the same
tile schedule, next-tile prefetch, per-tile counting and base fetch as the CUDA code, with the threads of a
CTA run in order and the CTAs interleaved round by round. They check the ALGORITHM (which tile's activity
bits a CTA uses when it walks several tiles, partial last tiles, counter re-zeroing) on grids small enough
for numpy -- the CUDA kernels themselves are checked on the GPU (tests/test_scatter_packet_gpu.py). The queue
model reproduces the defect fixed in the kernel (activity bits of a full tile reused for a partial last tile
when there is no mask array) when run with `refresh_without_mask=False`."""
import numpy as np
import pytest

from tests.test_scatter_packet_gpu import check_scatter_inc

THREADS = 256


def queue_model(size, grid, mask=None, refresh_without_mask=True, base0=7, rounds=8):
    """scatter_inc_queue_kernel: lane owns 4 consecutive elements per round, tile = THREADS * 4 * rounds"""
    tile = THREADS * 4 * rounds
    tiles = (size + tile - 1) // tile
    counter = base0
    out = np.full(size, -1, np.int64)
    lanes = np.arange(THREADS)

    def fetch(t):
        m = np.zeros((rounds, THREADS, 4), bool)
        for r in range(rounds):
            for k in range(4):
                e = t * tile + (r * THREADS + lanes) * 4 + k
                ok = e < size
                if mask is not None:
                    ok[ok] = mask[e[ok]] != 0
                m[r, :, k] = ok
        return m

    ctas = min(grid, tiles)
    regs = {c: fetch(c) for c in range(ctas)}          # `m_n` of every CTA
    it = 0
    while it * grid < tiles:
        for c in range(ctas):
            t = c + it * grid
            if t >= tiles:
                continue
            m = regs[c]
            if (mask is not None or refresh_without_mask) and t + grid < tiles:
                regs[c] = fetch(t + grid)
            cnt = m.sum(axis=(0, 2))
            base, counter = counter, counter + int(cnt.sum())          # the tile's one global atomic
            off = base + np.concatenate(([0], np.cumsum(cnt)[:-1]))
            for r in range(rounds):
                for k in range(4):
                    e = t * tile + (r * THREADS + lanes) * 4 + k
                    on = m[r, :, k]
                    inside = e < size
                    out[e[inside]] = np.where(on, off, 0)[inside]
                    off = off + on
        it += 1
    return counter, out


def private_model(size, grid, index, counters_before, mask=None, per_thread=8):
    """scatter_inc_private_kernel: element (j, tid) of a tile at base + j * THREADS + tid"""
    tile = THREADS * per_thread
    tiles = (size + tile - 1) // tile
    B = counters_before.size
    target = counters_before.astype(np.int64).copy()
    out = np.full(size, -1, np.int64)
    lanes = np.arange(THREADS)

    def fetch(t):
        idx = np.zeros((per_thread, THREADS), np.int64); ok = np.zeros((per_thread, THREADS), bool)
        for j in range(per_thread):
            i = t * tile + j * THREADS + lanes
            o = i < size
            if mask is not None:
                o[o] = mask[i[o]] != 0
            idx[j, o] = index[i[o]]
            ok[j] = o & (idx[j] < B)
        return idx, ok

    ctas = min(grid, tiles)
    regs = {c: fetch(c) for c in range(ctas)}
    scnt = {c: np.zeros(B, np.int64) for c in range(ctas)}
    it = 0
    while it * grid < tiles:
        for c in range(ctas):
            t = c + it * grid
            if t >= tiles:
                continue
            idx, ok = regs[c]
            if t + grid < tiles:
                regs[c] = fetch(t + grid)
            assert not scnt[c].any()                    # zero again after the previous tile
            loc = np.zeros((per_thread, THREADS), np.int64)
            for j in range(per_thread):
                for tid in np.flatnonzero(ok[j]):       # shared-memory atomics in some order
                    loc[j, tid] = scnt[c][idx[j, tid]]
                    scnt[c][idx[j, tid]] += 1
            sbase = np.full(B, -10 ** 9)                # stale entries must never be read
            for b in np.flatnonzero(scnt[c]):
                sbase[b] = target[b]; target[b] += scnt[c][b]; scnt[c][b] = 0
            for j in range(per_thread):
                i = t * tile + j * THREADS + lanes
                inside = i < size
                out[i[inside]] = np.where(ok[j], sbase[np.where(ok[j], idx[j], 0)] + loc[j], 0)[inside]   # (read only when active)
        it += 1
    return target, out


@pytest.mark.parametrize("size,grid", [(3 * 8192 + 77, 2), (5 * 8192 - 3, 2), (2 * 8192, 4), (100, 8), (8192 + 1, 1)])
def test_queue_model(size, grid):
    for mask in (None, np.random.RandomState(size).randint(0, 2, size).astype(np.uint8)):
        counter, out = queue_model(size, grid, mask)
        active = np.ones(size, bool) if mask is None else mask != 0
        assert counter == 7 + active.sum()
        assert np.array_equal(np.sort(out[active]), np.arange(7, 7 + active.sum())) and not out[~active].any()


def test_queue_model_reproduces_the_fixed_defect():
    """a CTA that walks from a full tile into the partial last tile without recomputing the bits"""
    size, grid = 3 * 8192 + 77, 2
    counter, _ = queue_model(size, grid, None, refresh_without_mask=False)
    assert counter != 7 + size
    counter, _ = queue_model(2 * 8192, 4, None, refresh_without_mask=False)      # one tile per CTA: not affected
    assert counter == 7 + 2 * 8192


@pytest.mark.parametrize("size,grid,B", [(3 * 2048 + 5, 2, 7), (2048, 4, 1), (5 * 2048 - 1, 2, 300), (17, 3, 2)])
def test_private_model(size, grid, B):
    rng = np.random.RandomState(size + B)
    index = rng.randint(0, B + 2, size).astype(np.uint32)           # two values past the counter array: ignored
    before = rng.randint(0, 50, B).astype(np.uint32)
    for mask in (None, rng.randint(0, 3, size) != 0):
        after, out = private_model(size, grid, index, before, None if mask is None else mask.astype(np.uint8))
        in_range = index < B
        eff_mask = in_range if mask is None else (mask & in_range)
        check_scatter_inc(before, after.astype(np.uint32), np.where(in_range, index, 0).astype(np.uint32), eff_mask,
                          out.astype(np.uint32), (size, grid, B))
