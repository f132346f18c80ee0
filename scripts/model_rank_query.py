#!/usr/bin/env python
"""numpy model of the rank-query copy-out of mkperm_tile_scatter16_kernel:
checks the algorithm -- boundary bitmap over the tile's slots, per-word prefix counts, rank -> bucket table,
bucket of slot j = segb[#starts <= j, minus 1] -- on tiles with empty buckets, one-bucket tiles and ragged tiles.
Not a test of the CUDA code (that is `DRJIT_B200_MKPERM_KPT=60 pytest -m gpu -k mkperm`)."""
import numpy as np


def tile_copy_out(keys, buckets, stride):
    n = keys.size
    cnt = np.bincount(keys, minlength=stride)
    start = np.cumsum(cnt) - cnt                                   # cursor[] before the ranking
    words = (n + 31) // 32
    bitmap = np.zeros(words + 1, np.uint32)
    for b in np.nonzero(cnt)[0]:                                   # phase (1): one bit per non-empty bucket start
        bitmap[start[b] >> 5] |= np.uint32(1) << np.uint32(start[b] & 31)
    pop = np.array([bin(int(w)).count("1") for w in bitmap])
    wrank = np.cumsum(pop) - pop                                   # phase (2)
    segb = np.full(stride, -1)
    for b in range(stride):                                        # phase (3)
        en = start[b + 1] if b + 1 < stride else n
        if en != start[b]:
            st = int(start[b])
            r = wrank[st >> 5] + bin(int(bitmap[st >> 5]) & ((1 << (st & 31)) - 1)).count("1")
            segb[r] = b
    cursor = start.copy()                                          # phase (4): ranking (any order inside a bucket)
    sorted16 = np.zeros(n, np.int64)
    for i in np.random.default_rng(1).permutation(n):
        sorted16[cursor[keys[i]]] = i
        cursor[keys[i]] += 1
    bucket_of_slot = np.empty(n, np.int64)                         # phase (5)
    for j in range(n):
        w, lane = j >> 5, j & 31
        le = (1 << (lane + 1)) - 1
        r = wrank[w] + bin(int(bitmap[w]) & le).count("1") - 1
        bucket_of_slot[j] = segb[r]
    assert np.array_equal(bucket_of_slot, keys[sorted16]), "slot -> bucket mapping is wrong"
    assert np.array_equal(np.sort(sorted16), np.arange(n))


def main():
    rng = np.random.default_rng(0)
    for n, buckets in [(2048, 64), (2048, 4096), (1000, 37), (33, 5), (1, 3), (4096, 1)]:
        stride = (buckets + 7) // 8 * 8
        tile_copy_out(rng.integers(0, buckets, n), buckets, stride)                     # uniform
        tile_copy_out(np.full(n, buckets - 1), buckets, stride)                         # one bucket holds everything
        tile_copy_out(rng.integers(0, buckets, n) // 7 * 7 % buckets, buckets, stride)  # many empty buckets
    print("rank-query model: ok")


if __name__ == "__main__":
    main()
