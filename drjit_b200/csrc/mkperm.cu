/*
 * mkperm.cu -- counting-sort permutation (vcall / dr.dispatch bucketing, dr.sort passes).
 *
 * Replaces CUDAThreadState::block_mkperm (ext/drjit-core/src/cuda_ts.cpp:788-975), cuda_transpose
 * (:765-786) and the kernels block_mkperm_phase_{1,3,4}_* / transpose (resources/mkperm.cuh:14-499).
 *
 * Pipeline (5 launches; the reference needs phase1 + transpose + memset + scan + transpose +
 * phase3 + memcpy + phase4):
 *   1. histogram  : every "row" (a warp or a CTA, see below) counts the keys of one contiguous
 *                   slice of a sorting group in shared memory, 128-bit key loads.
 *   2. column scan: per bucket, exclusive running sum over the rows of the group + bucket total
 *                   (replaces the two transposes around the reference's generic prefix sum).
 *   3. bucket scan: per group, exclusive scan of the bucket totals; for the single-group
 *                   (vcall) case also the table of non-empty buckets {id,start,size,0} in
 *                   ascending id order and the unique count, written straight to pinned memory.
 *   4. scatter    : every row re-reads its slice and writes the permutation.
 *
 * Variants (chosen by how many bucket counters fit into the 227 KB of shared memory):
 *   WARP   : one private histogram per warp -> the permutation is *stable* (identical to the
 *            reference's CPU backend, llvm_ts.cpp:785-933). Ranks inside a 32-key step come from
 *            match.any. Used while >= 8 warps fit (bucket_count <= 7264); the reference's
 *            stable "tiny" variant stops at 512 buckets per 64 KiB (cuda_ts.cpp:824-836).
 *   CTA    : one histogram per CTA, shared-memory atomics; valid but not stable across warps
 *            (same contract as the reference's "small" variant, jit.h:2404-2406).
 *   GLOBAL : global-memory atomics for bucket counts beyond shared memory ("large").
 */
#include <cooperative_groups.h>

#include "common.cuh"
#include "comm.cuh"
#include "tma.cuh"
#include "runtime.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace djb {

enum class MkpermMode : int { Warp = 0, Cta = 1, Global = 2 };

struct MkpermParams {
    const uint32_t *values;
    uint32_t *perm;
    uint32_t *rows;          // [group][row][bucket] histogram -> exclusive row offsets
    uint32_t *totals;        // [group][bucket] bucket totals -> exclusive bucket starts
    uint32_t size, block_size, bucket_count;
    uint32_t n_groups, ctas_per_group, rows_per_group, row_elems; // row = slice of a group
    uint32_t index_base;
    uint32_t row_stride;     // elements between histogram rows (0: bucket_count)
    uint32_t key_bits;       // ceil(log2(bucket_count))
    uint8_t vec;
};

/// Slice [start, end) of the group handled by row `row_in_group`
__device__ __forceinline__ void row_range(const MkpermParams &p, uint32_t group, uint32_t row_in_group,
                                          uint64_t &start, uint64_t &end) {
    const uint64_t group_start = (uint64_t) group * p.block_size;
    uint64_t group_end = group_start + p.block_size;
    if (group_end > p.size) group_end = p.size;
    start = group_start + (uint64_t) row_in_group * p.row_elems;
    end = start + p.row_elems;
    if (start > group_end) start = group_end;
    if (end > group_end) end = group_end;
}

// ---------------------------------------------------------------------------
//  Phase 1: histograms
// ---------------------------------------------------------------------------
template <MkpermMode Mode>
__global__ void mkperm_histogram_kernel(const MkpermParams p) {
    extern __shared__ uint32_t smem[];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u, nwarps = blockDim.x >> 5;
    const uint32_t group = blockIdx.x / p.ctas_per_group, cta = blockIdx.x - group * p.ctas_per_group;
    const uint32_t B = p.bucket_count;
    // Out-of-range keys (>= bucket_count; undefined behaviour in the reference) are counted in the
    // last bucket on every path of this file: one policy, no out-of-bounds counter access.
    const uint32_t last = B - 1;

    uint32_t *hist;         // counters this thread adds to
    uint32_t row_in_group;  // slice this warp reads
    if constexpr (Mode == MkpermMode::Warp) {
        hist = smem + (size_t) warp * B;
        row_in_group = cta * nwarps + warp;
        for (uint32_t i = lane; i < B; i += 32) hist[i] = 0;
        __syncwarp();
    } else if constexpr (Mode == MkpermMode::Cta) {
        hist = smem;
        row_in_group = cta * nwarps + warp;          // warps still read contiguous sub-slices
        for (uint32_t i = tid; i < B; i += blockDim.x) hist[i] = 0;
        __syncthreads();
    } else {
        hist = p.totals + (size_t) group * B;        // zeroed by the host
        row_in_group = cta * nwarps + warp;
    }

    // In Cta/Global mode `row_elems` is still the per-warp slice length
    uint64_t start, end;
    row_range(p, group, row_in_group, start, end);

    if (p.vec) {
        const uint64_t nvec = (end - start) / 4;
        const uint4 *v = reinterpret_cast<const uint4 *>(p.values + start);
        for (uint64_t i = lane; i < nvec; i += 4 * 32) {
            Vec16<uint32_t> t[4];
            bool ok[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                ok[u] = i + u * 32 < nvec;
                if (ok[u]) t[u] = ld_stream<uint32_t>(v + i + u * 32);
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ok[u]) {
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) atomicAdd(hist + min(t[u].v[e], last), 1u);
                }
        }
        for (uint64_t i = start + nvec * 4 + lane; i < end; i += 32)
            atomicAdd(hist + min(p.values[i], last), 1u);
    } else {
        for (uint64_t i = start + lane; i < end; i += 32)
            atomicAdd(hist + min(__ldg(p.values + i), last), 1u);
    }

    if constexpr (Mode == MkpermMode::Warp) {
        __syncwarp();
        uint32_t *dst = p.rows + ((size_t) group * p.rows_per_group + row_in_group) * B;
        for (uint32_t i = lane; i < B; i += 32) dst[i] = hist[i];
    } else if constexpr (Mode == MkpermMode::Cta) {
        __syncthreads();
        uint32_t *dst = p.rows + ((size_t) group * p.rows_per_group + cta) * B;
        for (uint32_t i = tid; i < B; i += blockDim.x) dst[i] = hist[i];
    }
}

// ---------------------------------------------------------------------------
//  Phase 2: per bucket, exclusive running sum over the rows of a group
// ---------------------------------------------------------------------------
/// One CTA = 32 adjacent buckets (lane = bucket, so every row access is one 128-byte line);
/// its 8 warps split the rows of the group into 8 contiguous segments: segment sums first,
/// then each warp rewrites its segment with the exclusive running values.
__global__ void __launch_bounds__(256)
mkperm_column_scan_kernel(const MkpermParams p, uint32_t tiles_per_group) {
    __shared__ uint32_t seg_sum[8][32];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t group = blockIdx.x / tiles_per_group,
                   b = (blockIdx.x - group * tiles_per_group) * 32 + lane;
    const bool valid = b < p.bucket_count;
    const uint32_t R = p.rows_per_group, seg = (R + 7) / 8,
                   r0 = min(warp * seg, R), r1 = min(r0 + seg, R);
    const size_t B = p.row_stride ? p.row_stride : p.bucket_count;   // elements between rows
    uint32_t *col = p.rows + (size_t) group * R * B + b;

    uint32_t sum = 0;
    if (valid) {
        uint32_t r = r0;
        for (; r + 8 <= r1; r += 8) {
            uint32_t t[8];
            #pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = col[(size_t) (r + u) * B];
            #pragma unroll
            for (int u = 0; u < 8; ++u) sum += t[u];
        }
        for (; r < r1; ++r) sum += col[(size_t) r * B];
    }
    seg_sum[warp][lane] = sum;
    __syncthreads();
    uint32_t running = 0, total = 0;
    #pragma unroll
    for (uint32_t w = 0; w < 8; ++w) {
        if (w == warp) running = total;
        total += seg_sum[w][lane];
    }
    if (!valid)
        return;
    uint32_t r = r0;
    for (; r + 8 <= r1; r += 8) {
        uint32_t t[8];
        #pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = col[(size_t) (r + u) * B];
        #pragma unroll
        for (int u = 0; u < 8; ++u) { col[(size_t) (r + u) * B] = running; running += t[u]; }
    }
    for (; r < r1; ++r) {
        const uint32_t t = col[(size_t) r * B];
        col[(size_t) r * B] = running;
        running += t;
    }
    if (warp == 0)
        p.totals[(size_t) group * p.bucket_count + b] = total;
}

// ---------------------------------------------------------------------------
//  Phase 3: per group, exclusive scan of bucket totals (+ table of non-empty buckets)
// ---------------------------------------------------------------------------
constexpr uint32_t kBucketScanThreads = 1024;

/// PEER (one sorting group sharded over the ranks of a communicator): the
/// kernel first exchanges the shard histograms through peer memory (comm.cuh). The table of
/// non-empty buckets then describes the GLOBAL array (start / size over all ranks, identical on
/// every rank), rank_base[b] is where this rank's keys of bucket b begin in the global, rank-major
/// stable order, while `totals` still receives the shard-local bucket starts the scatter pass needs.
///
/// sort_pow2 != 0 (jit_var_call_reduce, src/call.cpp:1346-1356: the dispatcher walks the buckets by
/// decreasing size): the rows are collected in shared memory, bitonic-sorted by (size descending, id
/// ascending) over sort_pow2 >= bucket_count slots and only then written to the table, so the host
/// needs no std::sort after the wait. Dynamic shared memory: sort_pow2 * 16 bytes.
template <bool PEER>
__global__ void __launch_bounds__(kBucketScanThreads)
mkperm_bucket_scan_kernel(const MkpermParams p, uint32_t *offsets, uint32_t *unique_out,
                          uint32_t *hist_out, const PeerCtx peer, uint32_t *rank_base, uint32_t sort_pow2) {
    extern __shared__ __align__(16) uint64_t sort_mem[];    // [sort_pow2] keys, then [sort_pow2] rows {id, start}
    uint64_t *sort_key = sort_mem;
    uint2 *sort_row = reinterpret_cast<uint2 *>(sort_mem + sort_pow2);
    __shared__ uint32_t warp_sum[32], warp_uniq[32], warp_gsum[32];
    __shared__ uint32_t carry_sum, carry_uniq, carry_gsum, epoch_smem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t group = blockIdx.x, B = p.bucket_count;
    uint32_t *totals = p.totals + (size_t) group * B;

    if (tid == 0) { carry_sum = 0; carry_uniq = 0; carry_gsum = 0; if (PEER) epoch_smem = peer_begin(peer); }
    for (uint32_t i = tid; i < sort_pow2; i += kBucketScanThreads) sort_key[i] = ~0ull;     // padding sorts last
    __syncthreads();
    uint32_t epoch = 0;
    if constexpr (PEER) {
        epoch = epoch_smem;
        peer_put_cta(peer, epoch, totals, B * 4);
        peer_wait_cta(peer, epoch);
    }

    for (uint32_t base = 0; base < B; base += kBucketScanThreads) {
        const uint32_t b = base + tid;
        const uint32_t n = b < B ? totals[b] : 0u;
        uint32_t gn = n, before = 0;        // bucket size over all ranks / over the ranks below mine
        if constexpr (PEER) {
            gn = 0;
            if (b < B) {
                for (uint32_t r = 0; r < peer.world; ++r) {
                    const uint32_t c = __ldcg(reinterpret_cast<const uint32_t *>(win_slot(peer, peer.rank, epoch, r)) + b);
                    if (r == peer.rank) before = gn;
                    gn += c;
                }
            }
        }
        const uint32_t u = gn != 0;
        if (hist_out && b < B) hist_out[b] = n;

        uint32_t vs = n, vu = u, vg = gn;
        #pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t ts = shfl_up(vs, d), tu = shfl_up(vu, d), tg = shfl_up(vg, d);
            if (lane >= d) { vs += ts; vu += tu; vg += tg; }
        }
        if (lane == 31) { warp_sum[warp] = vs; warp_uniq[warp] = vu; warp_gsum[warp] = vg; }
        __syncthreads();
        uint32_t ws = 0, wu = 0, wg = 0, ts_all = 0, tu_all = 0, tg_all = 0;
        #pragma unroll
        for (uint32_t w = 0; w < 32; ++w) {
            if (w == warp) { ws = ts_all; wu = tu_all; wg = tg_all; }
            ts_all += warp_sum[w]; tu_all += warp_uniq[w]; tg_all += warp_gsum[w];
        }
        const uint32_t start = carry_sum + ws + vs - n,   // exclusive, shard-local
                       gstart = PEER ? carry_gsum + wg + vg - gn : start,
                       slot = carry_uniq + wu + vu - u;
        if (b < B) {
            totals[b] = start;
            if (PEER && rank_base) rank_base[b] = gstart + before;
            if (offsets && u) {       // quadruple layout: jit.h:2412-2419, mkperm.cuh:271-320
                if (sort_pow2) {
                    sort_key[slot] = ((uint64_t) ~gn << 32) | slot;     // size descending, then id ascending
                    sort_row[slot] = make_uint2(b, gstart);
                } else {
                    uint4 q = make_uint4(b, gstart, gn, 0u);
                    *reinterpret_cast<uint4 *>(offsets + 4 * (size_t) slot) = q;
                }
            }
        }
        __syncthreads();
        if (tid == 0) { carry_sum += ts_all; carry_uniq += tu_all; carry_gsum += tg_all; }
        __syncthreads();
    }
    if (sort_pow2 && offsets) {
        for (uint32_t k = 2; k <= sort_pow2; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = tid; i < sort_pow2; i += kBucketScanThreads) {
                    const uint32_t partner = i ^ j;
                    if (partner > i) {
                        const uint64_t a = sort_key[i], c = sort_key[partner];
                        if ((a > c) == ((i & k) == 0)) { sort_key[i] = c; sort_key[partner] = a; }
                    }
                }
                __syncthreads();
            }
        }
        for (uint32_t i = tid; i < carry_uniq; i += kBucketScanThreads) {
            const uint64_t key = sort_key[i];
            const uint2 row = sort_row[(uint32_t) key];
            *reinterpret_cast<uint4 *>(offsets + 4 * (size_t) i) = make_uint4(row.x, row.y, ~(uint32_t) (key >> 32), 0u);
        }
    }
    if (tid == 0 && offsets) {
        offsets[4 * (size_t) B] = carry_uniq;  // cuda_ts.cpp:948-951
        if (unique_out) *unique_out = carry_uniq;
    }
    if (PEER && tid == 0)
        peer_end(peer, epoch);
}

/// Sharded call: communicator view + optional output of this rank's start inside every global bucket
struct MkpermPeer { PeerCtx ctx; uint32_t *rank_base; };

constexpr uint32_t kMkpermMaxPayloads = 4;
constexpr uint32_t kSortTableMaxBuckets = 8192;     // device-side sort of the bucket table (128 KiB of shared memory)

/// Extras of jit_var_call_reduce (src/call.cpp:1268-1389) on top of the plain block_mkperm:
/// the table sorted by decreasing bucket size, and 32-bit payload arrays permuted along with `perm`
struct MkpermExtras {
    bool sort_table = false;
    uint32_t n_pay = 0;
    const uint32_t *pay_in[kMkpermMaxPayloads] = {};
    uint32_t *pay_out[kMkpermMaxPayloads] = {};
};

/// Launches the bucket scan of `p` (PEER when a communicator is given)
static void launch_bucket_scan(cudaStream_t stream, const MkpermParams &p, uint32_t *offsets_dev, uint32_t *unique_dev,
                               uint32_t *hist_out, const MkpermPeer *peer, bool sort_table = false) {
    uint32_t sort_pow2 = 0, smem = 0;
    if (sort_table && offsets_dev && p.n_groups == 1) {
        sort_pow2 = std::max(2u, round_pow2(p.bucket_count));
        smem = sort_pow2 * 16;
        static std::atomic<bool> configured_on[kMaxDevices] = {};
        std::atomic<bool> &configured = configured_on[device_props().device % kMaxDevices];
        if (!configured.load(std::memory_order_acquire)) {
            DJB_CUDA_CHECK(cudaFuncSetAttribute(mkperm_bucket_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kSortTableMaxBuckets * 16)));
            DJB_CUDA_CHECK(cudaFuncSetAttribute(mkperm_bucket_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kSortTableMaxBuckets * 16)));
            configured.store(true, std::memory_order_release);
        }
    }
    if (peer)
        mkperm_bucket_scan_kernel<true><<<p.n_groups, kBucketScanThreads, smem, stream>>>(p, offsets_dev, unique_dev, hist_out, peer->ctx, peer->rank_base, sort_pow2);
    else
        mkperm_bucket_scan_kernel<false><<<p.n_groups, kBucketScanThreads, smem, stream>>>(p, offsets_dev, unique_dev, hist_out, PeerCtx{}, nullptr, sort_pow2);
    DJB_POST_LAUNCH();
}

// ---------------------------------------------------------------------------
//  Phase 4: scatter
// ---------------------------------------------------------------------------
template <MkpermMode Mode>
__global__ void mkperm_scatter_kernel(const MkpermParams p) {
    extern __shared__ uint32_t smem[];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u, nwarps = blockDim.x >> 5;
    const uint32_t group = blockIdx.x / p.ctas_per_group, cta = blockIdx.x - group * p.ctas_per_group;
    const uint32_t B = p.bucket_count, last = B - 1;    // (out-of-range keys: last bucket, as in the histogram)
    const uint32_t group_start = group * p.block_size;   // < size, fits
    const uint32_t *bucket_start = p.totals + (size_t) group * B;
    const uint32_t row_in_group = cta * nwarps + warp;

    uint32_t *ctr;
    if constexpr (Mode == MkpermMode::Warp) {
        ctr = smem + (size_t) warp * B;
        const uint32_t *src = p.rows + ((size_t) group * p.rows_per_group + row_in_group) * B;
        for (uint32_t i = lane; i < B; i += 32) ctr[i] = group_start + bucket_start[i] + src[i];
        __syncwarp();
    } else if constexpr (Mode == MkpermMode::Cta) {
        ctr = smem;
        const uint32_t *src = p.rows + ((size_t) group * p.rows_per_group + cta) * B;
        for (uint32_t i = tid; i < B; i += blockDim.x) ctr[i] = group_start + bucket_start[i] + src[i];
        __syncthreads();
    } else {
        ctr = p.totals + (size_t) group * B; // running global cursors (group-relative)
    }

    uint64_t start, end;
    row_range(p, group, row_in_group, start, end);

    if constexpr (Mode == MkpermMode::Warp) {
        // Stable: 32 consecutive keys per step, ranks among equal keys from match.any
        for (uint64_t base = start; base < end; base += 4 * 32) {
            uint32_t key[4];
            bool ok[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t i = base + u * 32 + lane;
                ok[u] = i < end;
                key[u] = ok[u] ? min(__ldg(p.values + i), last) : 0u;
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t active = __ballot_sync(kFullMask, ok[u]);
                if (active == 0) break;
                // lanes holding the same key: one ballot per key bit (a match.any occupies a unit
                // that is shared by the whole SM for ~64 cycles, scripts/microbench.cu)
                uint32_t pos = 0, peers = active;
                for (uint32_t bit = 0; bit < p.key_bits; ++bit) {
                    const bool one = ok[u] && ((key[u] >> bit) & 1u);
                    const uint32_t v = __ballot_sync(kFullMask, one);
                    peers &= one ? v : ~v;
                }
                if (ok[u])
                    pos = ctr[key[u]] + __popc(peers & lanemask_lt());
                __syncwarp();
                if (ok[u] && (peers & lanemask_lt()) == 0)      // lowest lane of the peer group
                    ctr[key[u]] += __popc(peers);
                __syncwarp();
                if (ok[u])
                    p.perm[pos] = p.index_base + (uint32_t) (base + u * 32 + lane);
            }
        }
    } else {
        auto place = [&](uint32_t key, uint32_t i) {
            uint32_t pos = atomicAdd(ctr + min(key, last), 1u);
            if constexpr (Mode == MkpermMode::Global) pos += group_start;
            p.perm[pos] = p.index_base + i;
        };
        if (p.vec) {
            const uint64_t nvec = (end - start) / 4;
            const uint4 *v = reinterpret_cast<const uint4 *>(p.values + start);
            for (uint64_t i = lane; i < nvec; i += 4 * 32) {
                Vec16<uint32_t> t[4];
                bool ok[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ok[u] = i + u * 32 < nvec;
                    if (ok[u]) t[u] = ld_stream<uint32_t>(v + i + u * 32);
                }
                #pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ok[u]) {
                        #pragma unroll
                        for (int e = 0; e < 4; ++e)
                            place(t[u].v[e], (uint32_t) (start + (i + u * 32) * 4 + e));
                    }
            }
            for (uint64_t i = start + nvec * 4 + lane; i < end; i += 32)
                place(p.values[i], (uint32_t) i);
        } else {
            for (uint64_t i = start + lane; i < end; i += 32)
                place(__ldg(p.values + i), (uint32_t) i);
        }
    }
}

// ---------------------------------------------------------------------------
//  Tile-staged fast path (single sorting group, bucket_count <= 8192, large inputs)
// ---------------------------------------------------------------------------
//  The row-per-warp design above keeps one write cursor per (warp, bucket): with 4096 buckets
//  and ~1900 rows that is 7.9 M cursors advancing 4 bytes at a time, the partially written
//  sectors do not survive in L2 and every 4-byte store costs a 32-byte DRAM read-modify-write
//  (measured: 1.93 GB written + 1.96 GB read for a 268 MB permutation). Here the array is cut
//  into tiles that all CTAs walk in the same order, so that at any time the whole chip writes
//  into a narrow moving window of every bucket (a few KiB) that L2 merges into full lines:
//    K1 tile histogram : CTA c counts a contiguous chunk of tiles in shared memory (8.2 shared
//                        atomics/clk/SM measured); per tile it records the running counts
//                        before the tile (tile_off, exclusive prefix inside the chunk) and
//                        the tile's own counts (cnt16).
//    K2 column scan over the chunk totals + bucket scan (kernels above).
//    K3 tile scatter   : per tile: (1) bins: exclusive scan of the tile's counts -> one write
//                        cursor per bucket in shared memory, plus the distance between a slot of
//                        the tile-local order and its final position (all 128-bit loads/stores,
//                        issued while the keys are still in flight); (2) keys: slot = shared
//                        atomicAdd on the cursor, entry (bucket, local index) stored at the slot:
//                        two shared-memory operations per key; (3) slots are copied out in order:
//                        runs of equal buckets are contiguous in shared memory and in `perm`, so
//                        consecutive threads store consecutive entries.
//  The kernel is bound by the LSU pipe -- shared-memory wavefronts plus the sectors of the run-wise
//  global stores (profiles/r1b: 11.5 sectors per store request at 16 Ki keys per tile and 4096
//  buckets) -- which is why tiles are as large as shared memory allows: runs get longer.
//  Ranks come from atomics, so the order inside a bucket is not the input order (same contract
//  as the reference's "small"/"large" variants, jit.h:2404-2406); bucket boundaries, the offsets
//  table and per-bucket contents are exact.
constexpr uint32_t kTileKeysPerThread = 32;
constexpr uint32_t kMkpermDefaultKpt = 48;  // largest tile of the unordered kernel, keys per thread (see mkperm_impl)
constexpr uint32_t kTileMaxBuckets = 8192;

struct MkpermTileParams {
    const uint32_t *values;
    uint32_t *perm;
    uint32_t *tile_off;      // [tiles][stride] exclusive prefix of the tile inside its chunk
    uint16_t *tile_cnt;      // [tiles][stride] counts of the tile
    uint32_t *rows;          // [chunks][stride] chunk totals -> exclusive chunk offsets
    const uint32_t *bucket_start; // [buckets] after the bucket scan
    uint32_t size, bucket_count, stride, tiles, tiles_per_chunk, index_base;
    // jit_var_call_reduce: 32-bit payload arrays that travel with the permutation,
    // pay_out[k][j] = pay_in[k][perm[j] - index_base] (the tile was just read: L2 hits)
    uint32_t n_pay;
    const uint32_t *pay_in[4];
    uint32_t *pay_out[4];
    uint8_t vec;
    uint8_t debug;           // read only in -DDRJIT_B200_EXPERIMENTS builds (phases switched off for timing):
                             // 1 = copy-out without its global stores, 2 = no copy-out, 4 = no ranking, 8 = no bucket rows
};

/// Loads the keys of one tile into registers (clamped to the last bucket: out-of-range keys are
/// undefined behaviour in the reference; here they can at least not corrupt shared memory).
/// local index of key (k, e): VEC: ((k * threads + tid) * 4 + e), else k * threads + tid.
template <uint32_t THREADS>
__device__ __forceinline__ void tile_load_keys(const MkpermTileParams &p, uint64_t tile_base, uint32_t n_tile,
                                               uint32_t (&key)[kTileKeysPerThread]) {
    const uint32_t tid = threadIdx.x, last = p.bucket_count - 1;
    if (p.vec && n_tile == THREADS * kTileKeysPerThread) {
        const uint4 *v = reinterpret_cast<const uint4 *>(p.values + tile_base);
        #pragma unroll
        for (uint32_t k = 0; k < kTileKeysPerThread / 4; ++k) {
            const Vec16<uint32_t> t = ld_stream<uint32_t>(v + k * THREADS + tid);
            #pragma unroll
            for (uint32_t e = 0; e < 4; ++e) key[k * 4 + e] = min(t.v[e], last);
        }
    } else {
        #pragma unroll
        for (uint32_t k = 0; k < kTileKeysPerThread; ++k) {
            const uint32_t i = k * THREADS + tid;
            key[k] = i < n_tile ? min(__ldg(p.values + tile_base + i), last) : 0xffffffffu;
        }
    }
}

/// Same for tiles of more than 32 keys per thread: two keys per register (bucket ids on the tile
/// path are < 8192; 0xffff marks a slot past the end), key k = half (k & 1) of kp[k >> 1], same
/// local indices as above.
template <uint32_t THREADS, uint32_t KPT>
__device__ __forceinline__ void tile_load_keys_packed(const MkpermTileParams &p, uint64_t tile_base, uint32_t n_tile,
                                                      uint32_t (&kp)[KPT / 2]) {
    static_assert(KPT % 4 == 0, "keys are loaded four at a time");
    const uint32_t tid = threadIdx.x, last = p.bucket_count - 1;
    if (p.vec && n_tile == THREADS * KPT) {
        const uint4 *v = reinterpret_cast<const uint4 *>(p.values + tile_base);
        #pragma unroll
        for (uint32_t k = 0; k < KPT / 4; ++k) {
            const Vec16<uint32_t> t = ld_stream<uint32_t>(v + k * THREADS + tid);
            kp[2 * k] = min(t.v[0], last) | (min(t.v[1], last) << 16);
            kp[2 * k + 1] = min(t.v[2], last) | (min(t.v[3], last) << 16);
        }
    } else {
        #pragma unroll
        for (uint32_t k = 0; k < KPT / 2; ++k) {
            const uint32_t i0 = (2 * k) * THREADS + tid, i1 = i0 + THREADS;
            const uint32_t a = i0 < n_tile ? min(__ldg(p.values + tile_base + i0), last) : 0xffffu,
                           b = i1 < n_tile ? min(__ldg(p.values + tile_base + i1), last) : 0xffffu;
            kp[k] = a | (b << 16);
        }
    }
}

/// Copy-out of a ranked tile together with the payload arrays of jit_var_call_reduce: the entry of
/// slot j goes to perm[delta[bucket] + j] and every payload value of that key follows it to the same
/// position. The payload tile is contiguous (tile_base ..) and was asked into L2 one tile ahead, so the
/// reads are L2 hits in 32-byte sectors that the CTA uses up completely; four slots per thread are in
/// flight at a time.
template <uint32_t THREADS>
__device__ __forceinline__ void tile_copy_out_payloads(const MkpermTileParams &p, const uint32_t *sorted, const uint32_t *delta,
                                                       uint64_t tile_base, uint32_t n_tile, uint32_t idx0) {
    for (uint32_t j0 = threadIdx.x; j0 < n_tile; j0 += 4 * THREADS) {
        uint32_t at[4], local[4];
        bool ok[4];
        #pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {
            const uint32_t j = j0 + u * THREADS;
            ok[u] = j < n_tile;
            const uint32_t e = ok[u] ? sorted[j] : 0u;
            at[u] = delta[e >> 16] + j; local[u] = e & 0xffffu;
        }
        #pragma unroll
        for (uint32_t u = 0; u < 4; ++u)
            if (ok[u]) p.perm[at[u]] = idx0 + local[u];
        for (uint32_t k = 0; k < p.n_pay; ++k) {
            const uint32_t *src = p.pay_in[k] + tile_base;
            uint32_t *dst = p.pay_out[k];
            uint32_t v[4];
            #pragma unroll
            for (uint32_t u = 0; u < 4; ++u) v[u] = ok[u] ? __ldg(src + local[u]) : 0u;
            #pragma unroll
            for (uint32_t u = 0; u < 4; ++u)
                if (ok[u]) dst[at[u]] = v[u];
        }
    }
}

/// Asks L2 for the payload tiles of tile `tile` (whole tiles of 16-byte aligned arrays only)
__device__ __forceinline__ void tile_prefetch_payloads(const MkpermTileParams &p, uint64_t tile, uint32_t tile_keys) {
    if ((tile + 1) * tile_keys > p.size)
        return;
    for (uint32_t k = 0; k < p.n_pay; ++k)
        if ((((uintptr_t) p.pay_in[k]) & 15u) == 0)
            bulk_prefetch_l2(p.pay_in[k] + tile * tile_keys, tile_keys * 4);
}

template <uint32_t THREADS, uint32_t KPT = kTileKeysPerThread>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
mkperm_tile_hist_kernel(const MkpermTileParams p) {
    constexpr uint32_t TILE = THREADS * KPT;
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t S = p.stride;
    uint32_t *hist = smem;                                  // running counts of this chunk
    uint32_t *prev = smem + S;                              // ... as of the previous tile
    const uint32_t tid = threadIdx.x;
    for (uint32_t b = tid; b < S; b += THREADS) { hist[b] = 0; prev[b] = 0; }

    const uint32_t first = blockIdx.x * p.tiles_per_chunk,
                   end = min(first + p.tiles_per_chunk, p.tiles);
    // snapshot before tile `tile` (tile == end: only the counts of the last tile are due)
    auto snapshot = [&](uint32_t tile) {
        for (uint32_t b = tid * 4; b < S; b += THREADS * 4) {
            const uint4 h = *reinterpret_cast<const uint4 *>(hist + b);
            const uint4 q = *reinterpret_cast<const uint4 *>(prev + b);
            if (tile < end)
                *reinterpret_cast<uint4 *>(p.tile_off + (size_t) tile * S + b) = h;
            if (tile > first) {
                const uint2 c = make_uint2((h.x - q.x) | ((h.y - q.y) << 16), (h.z - q.z) | ((h.w - q.w) << 16));
                *reinterpret_cast<uint2 *>(p.tile_cnt + (size_t) (tile - 1) * S + b) = c;
            }
            *reinterpret_cast<uint4 *>(prev + b) = h;
        }
    };
    for (uint32_t tile = first; tile < end; ++tile) {
        const uint64_t tile_base = (uint64_t) tile * TILE;
        const uint32_t n_tile = (uint32_t) min((uint64_t) TILE, (uint64_t) p.size - tile_base);
        constexpr bool PACKED = KPT != kTileKeysPerThread;  // (the unpacked loader is written for 32 keys per thread)
        uint32_t key[PACKED ? KPT / 2 : KPT];
        if constexpr (PACKED)
            tile_load_keys_packed<THREADS, KPT>(p, tile_base, n_tile, key);
        else
            tile_load_keys<THREADS>(p, tile_base, n_tile, key);  // loads in flight across the barrier
        if (tid == 0 && p.vec && tile + 1 < end && (uint64_t) (tile + 2) * TILE <= p.size)
            bulk_prefetch_l2(p.values + (uint64_t) (tile + 1) * TILE, TILE * 4);
        __syncthreads();                                    // previous tile's atomics are done
        snapshot(tile);
        __syncthreads();
        #pragma unroll
        for (uint32_t k = 0; k < KPT; ++k) {
            if constexpr (PACKED) {
                const uint32_t b = (key[k >> 1] >> (16 * (k & 1u))) & 0xffffu;
                if (b != 0xffffu) atomicAdd(hist + b, 1u);
            } else {
                if (key[k] != 0xffffffffu) atomicAdd(hist + key[k], 1u);
            }
        }
    }
    __syncthreads();
    snapshot(end);
    uint32_t *row = p.rows + (size_t) blockIdx.x * S;
    for (uint32_t b = tid; b < S; b += THREADS) row[b] = hist[b];
}

/// PAYS (jit_var_call_reduce with argument arrays, large inputs): the tile is half as large and the
/// other half of shared memory holds the tile of ONE payload array at a time: coalesced 128-bit loads
/// bring it in, the copy-out reads it at the staged local index (a shared-memory access instead of a
/// scattered 4-byte global load per key: those ran at one 32-byte sector per lane and instruction,
/// ~1.1 ms per array at 2^26 keys, as slow as a separate gather) and writes it where `perm` goes.
template <uint32_t THREADS, uint32_t KPT = kTileKeysPerThread, bool PAYS = false>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
mkperm_tile_scatter_kernel(const MkpermTileParams p) {
    constexpr uint32_t TILE = THREADS * KPT, WARPS = THREADS / 32;
    constexpr bool PACKED = KPT != kTileKeysPerThread;
    static_assert(TILE <= 65536, "local indices are packed into 16 bits");
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t S = p.stride;
    uint32_t *cursor = smem;             // [S] next free slot of the bucket in the tile-local order
    uint32_t *delta = smem + S;          // [S] final position of the bucket's run minus its local start
    uint32_t *sorted = smem + 2 * S;     // [TILE] (bucket << 16 | local index), ordered by bucket
    __shared__ uint32_t warp_sum[WARPS];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;

    // Tile order: round-robin over the CTAs (all CTAs advance through the input together, so the writes
    // of the chip stay inside a narrow moving window per bucket). Experiments builds can switch to a
    // blocked order (CTA c owns a contiguous range of tiles: the runs of successive tiles of a bucket
    // are adjacent in `perm` and come from the same SM) for A/B timing: DRJIT_B200_MKPERM_DEBUG bit 16.
    const bool blocked = DJB_DEBUG(p.debug) & 16u;
    const uint32_t per_cta = (p.tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t tile_first = blocked ? blockIdx.x * per_cta : blockIdx.x,
                   tile_end = blocked ? min(tile_first + per_cta, p.tiles) : p.tiles,
                   tile_step = blocked ? 1u : gridDim.x;
    for (uint32_t tile = tile_first; tile < tile_end; tile += tile_step) {
        const uint64_t tile_base = (uint64_t) tile * TILE;
        const uint32_t n_tile = (uint32_t) min((uint64_t) TILE, (uint64_t) p.size - tile_base);
        const bool vec = p.vec && n_tile == TILE;

        uint32_t key[PACKED ? KPT / 2 : KPT];
        if constexpr (PACKED)
            tile_load_keys_packed<THREADS, KPT>(p, tile_base, n_tile, key);
        else
            tile_load_keys<THREADS>(p, tile_base, n_tile, key); // in flight during the bin phase
        // Ask L2 for everything the next tile of this CTA will read, so that its loads see an L2
        // round trip instead of DRAM latency (the kernel was latency-bound: profiles/r1d)
        if (tid == 0) {
            const uint64_t next = (uint64_t) tile + tile_step;
            if (p.n_pay && tile == tile_first) tile_prefetch_payloads(p, tile, TILE);
            if (next < tile_end) {
                if (p.vec && (next + 1) * TILE <= p.size)
                    bulk_prefetch_l2(p.values + next * TILE, TILE * 4);
                bulk_prefetch_l2(p.tile_off + next * S, S * 4);
                bulk_prefetch_l2(p.tile_cnt + next * S, S * 2);
                if (p.n_pay) tile_prefetch_payloads(p, next, TILE);
            }
        }

        // ---- (1) bins: thread t owns 8 consecutive buckets per round ----------------------
        if (!(DJB_DEBUG(p.debug) & 8u)) {
            const uint32_t chunk = tile / p.tiles_per_chunk;
            const uint16_t *cnt = p.tile_cnt + (size_t) tile * S;
            const uint32_t *toff = p.tile_off + (size_t) tile * S, *crow = p.rows + (size_t) chunk * S;
            uint32_t carry = 0;
            for (uint32_t base = 0; base < S; base += THREADS * 8) {
                const uint32_t b0 = base + tid * 8;
                uint32_t c[8], g[8];
                #pragma unroll
                for (uint32_t j = 0; j < 8; ++j) { c[j] = 0; g[j] = 0; }
                if (b0 < S) {
                    const uint4 cc = __ldg(reinterpret_cast<const uint4 *>(cnt + b0));
                    const uint4 t0 = __ldg(reinterpret_cast<const uint4 *>(toff + b0)),
                                t1 = __ldg(reinterpret_cast<const uint4 *>(toff + b0 + 4)),
                                r0 = __ldg(reinterpret_cast<const uint4 *>(crow + b0)),
                                r1 = __ldg(reinterpret_cast<const uint4 *>(crow + b0 + 4));
                    uint32_t s8[8];     // bucket starts (the array has bucket_count entries: no vector loads past it)
                    #pragma unroll
                    for (uint32_t j = 0; j < 8; ++j) s8[j] = b0 + j < p.bucket_count ? __ldg(p.bucket_start + b0 + j) : 0u;
                    c[0] = cc.x & 0xffffu; c[1] = cc.x >> 16; c[2] = cc.y & 0xffffu; c[3] = cc.y >> 16;
                    c[4] = cc.z & 0xffffu; c[5] = cc.z >> 16; c[6] = cc.w & 0xffffu; c[7] = cc.w >> 16;
                    g[0] = t0.x + r0.x + s8[0]; g[1] = t0.y + r0.y + s8[1]; g[2] = t0.z + r0.z + s8[2]; g[3] = t0.w + r0.w + s8[3];
                    g[4] = t1.x + r1.x + s8[4]; g[5] = t1.y + r1.y + s8[5]; g[6] = t1.z + r1.z + s8[6]; g[7] = t1.w + r1.w + s8[7];
                }
                uint32_t sum = 0;
                #pragma unroll
                for (uint32_t j = 0; j < 8; ++j) sum += c[j];
                uint32_t incl = sum;
                #pragma unroll
                for (uint32_t d = 1; d < 32; d <<= 1) {
                    const uint32_t t = shfl_up(incl, d);
                    if (lane >= d) incl += t;
                }
                if (lane == 31) warp_sum[warp] = incl;
                __syncthreads();
                uint32_t wbase = 0, total = 0;
                #pragma unroll
                for (uint32_t w = 0; w < WARPS; ++w) {
                    if (w == warp) wbase = total;
                    total += warp_sum[w];
                }
                uint32_t run = carry + wbase + incl - sum;
                carry += total;
                if (b0 < S) {
                    uint32_t st[8];
                    #pragma unroll
                    for (uint32_t j = 0; j < 8; ++j) { st[j] = run; g[j] -= run; run += c[j]; }
                    *reinterpret_cast<uint4 *>(cursor + b0) = make_uint4(st[0], st[1], st[2], st[3]);
                    *reinterpret_cast<uint4 *>(cursor + b0 + 4) = make_uint4(st[4], st[5], st[6], st[7]);
                    *reinterpret_cast<uint4 *>(delta + b0) = make_uint4(g[0], g[1], g[2], g[3]);
                    *reinterpret_cast<uint4 *>(delta + b0 + 4) = make_uint4(g[4], g[5], g[6], g[7]);
                }
                __syncthreads();
            }
        }

        // ---- (2) keys: slot from the bucket's cursor, entry stored at the slot ---------------
        if (!(DJB_DEBUG(p.debug) & 4u)) {
            #pragma unroll
            for (uint32_t k = 0; k < KPT; ++k) {
                const uint32_t local = vec ? ((k / 4) * THREADS + tid) * 4 + (k & 3u) : k * THREADS + tid;
                if constexpr (PACKED) {
                    const uint32_t b = (key[k >> 1] >> (16 * (k & 1u))) & 0xffffu;
                    if (b != 0xffffu)
                        sorted[atomicAdd(cursor + b, 1u)] = (b << 16) | local;
                } else {
                    if (key[k] != 0xffffffffu)
                        sorted[atomicAdd(cursor + key[k], 1u)] = (key[k] << 16) | local;
                }
            }
        } else {
            uint32_t x = 0;                             // (keeps the key loads alive)
            #pragma unroll
            for (uint32_t k = 0; k < (PACKED ? KPT / 2 : KPT); ++k) x ^= key[k];
            if (x == 0x12345678u) sorted[tid] = x;
        }
        __syncthreads();

        // ---- (3) runs of equal buckets are contiguous in `sorted` and in `perm` ----------------
        const uint32_t idx0 = p.index_base + (uint32_t) tile_base;
        if (!(DJB_DEBUG(p.debug) & 15u) && p.n_pay && !PAYS) {
            tile_copy_out_payloads<THREADS>(p, sorted, delta, tile_base, n_tile, idx0);
        } else if (!(DJB_DEBUG(p.debug) & 15u)) {
            #pragma unroll 4
            for (uint32_t j = tid; j < n_tile; j += THREADS) {
                const uint32_t e = sorted[j];
                p.perm[delta[e >> 16] + j] = idx0 + (e & 0xffffu);
            }
            if constexpr (PAYS) {
                uint32_t *pay_tile = sorted + TILE;         // [TILE] one payload array's values of this tile
                for (uint32_t k = 0; k < p.n_pay; ++k) {
                    const uint32_t *src = p.pay_in[k] + tile_base;
                    uint32_t *dst = p.pay_out[k];
                    if (n_tile == TILE && (((uintptr_t) src) & 15u) == 0) {
                        #pragma unroll 4
                        for (uint32_t i = tid; i < TILE / 4; i += THREADS) {
                            const Vec16<uint32_t> t = ld_stream<uint32_t>(reinterpret_cast<const uint4 *>(src) + i);
                            *reinterpret_cast<uint4 *>(pay_tile + 4 * i) = *reinterpret_cast<const uint4 *>(&t);
                        }
                    } else {
                        for (uint32_t i = tid; i < n_tile; i += THREADS) pay_tile[i] = __ldg(src + i);
                    }
                    __syncthreads();
                    #pragma unroll 4
                    for (uint32_t j = tid; j < n_tile; j += THREADS) {
                        const uint32_t e = sorted[j];
                        dst[delta[e >> 16] + j] = pay_tile[e & 0xffffu];
                    }
                    __syncthreads();                        // the next array (or tile) overwrites pay_tile
                }
            }
        } else if ((DJB_DEBUG(p.debug) & 15u) == 1u) {
            #pragma unroll 4
            for (uint32_t j = tid; j < n_tile; j += THREADS) {
                const uint32_t e = sorted[j];
                const uint32_t at = delta[e >> 16] + j;
                if (at == 0xfffffff1u) p.perm[0] = idx0 + (e & 0xffffu);
            }
        }
        __syncthreads();
    }
}

#if defined(DRJIT_B200_EXPERIMENTS)
// ---------------------------------------------------------------------------
//  Unordered tile scatter, two tiles per cluster of two CTAs (runs merged through distributed shared memory)
// ---------------------------------------------------------------------------
//  What bounds the kernel above is the store path of its copy-out: a bucket's run is ~12 entries per
//  48 Ki-key tile at 4096 buckets, i.e. partial, unaligned lines towards L2. Here two CTAs on two SMs
//  (a thread-block cluster) rank two CONSECUTIVE tiles; the runs of a bucket in consecutive tiles are
//  adjacent in `perm`, so after a cluster barrier CTA r copies out the buckets of its half of the
//  bucket range for BOTH tiles -- its own staged entries from local shared memory, the other tile's
//  through distributed shared memory -- one warp per bucket at a time: lanes [0, cA) write tile A's
//  run, lanes [cA, cA + cB) tile B's, one store instruction covers both (runs of ~24 entries).
//  Measured in isolation (scripts/microbench_cluster.cu): 0.129 ms for the stores of 2^26 entries
//  instead of 0.228 ms. Phases (1) and (2) are those of the kernel above.
//  RESULT (profiles/r4k_*): correct (parity tests + racecheck) but SLOWER -- 0.56-0.62 ms against 0.35 ms
//  for the whole mkperm. The isolated measurement mapped merged slots to buckets with pure arithmetic
//  (uniform runs); real runs vary, and finding "which bucket, which tile" per store costs ~45
//  instructions and a remote shared-memory round trip per bucket where the single-tile copy-out
//  spends 6 instructions per 32 entries. EXPERIMENTS builds only (DRJIT_B200_MKPERM_PAIR=1).
template <uint32_t THREADS, uint32_t KPT>
__global__ void __launch_bounds__(THREADS, 1)
mkperm_tile_scatter_pair_kernel(const MkpermTileParams p) {
    namespace cg = cooperative_groups;
    constexpr uint32_t TILE = THREADS * KPT, WARPS = THREADS / 32;
    constexpr bool PACKED = KPT > kTileKeysPerThread;
    static_assert(TILE <= 65536, "local indices and run descriptors are packed into 16 bits");
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t S = p.stride;
    uint32_t *cursor = smem;             // [S] next free slot of the bucket -> (after ranking) end of its run
    uint32_t *delta = smem + S;          // [S] final position of the bucket's run minus its local start
    uint32_t *sorted = smem + 2 * S;     // [TILE] (bucket << 16 | local index), ordered by bucket
    __shared__ uint32_t warp_sum[WARPS];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();                     // 0: even tile of the pair, 1: odd tile
    const uint32_t pairs = (p.tiles + 1) / 2, pair_step = gridDim.x / 2;
    const uint32_t *tile_smem[2] = { rank == 0 ? smem : cluster.map_shared_rank(smem, 0),
                                     rank == 1 ? smem : cluster.map_shared_rank(smem, 1) };
    // my half of the buckets (split at a multiple of 32: a warp takes 32 buckets per round)
    const uint32_t split = min(S, ((S / 2 + 31) / 32) * 32), h0 = rank ? split : 0u, h1 = rank ? S : split;

    for (uint32_t pair = blockIdx.x / 2; pair < pairs; pair += pair_step) {
        const uint32_t tile = 2 * pair + rank;
        const bool present = tile < p.tiles;                        // (an odd tile count leaves the last pair half empty)
        const uint64_t tile_base = (uint64_t) tile * TILE;
        const uint32_t n_tile = present ? (uint32_t) min((uint64_t) TILE, (uint64_t) p.size - tile_base) : 0u;
        const bool vec = p.vec && n_tile == TILE;

        if (present) {
            uint32_t key[PACKED ? KPT / 2 : KPT];
            if constexpr (PACKED)
                tile_load_keys_packed<THREADS, KPT>(p, tile_base, n_tile, key);
            else
                tile_load_keys<THREADS>(p, tile_base, n_tile, key);
            if (tid == 0) {
                const uint64_t next = (uint64_t) tile + 2ull * pair_step;
                if (next < p.tiles) {
                    if (p.vec && (next + 1) * TILE <= p.size)
                        bulk_prefetch_l2(p.values + next * TILE, TILE * 4);
                    bulk_prefetch_l2(p.tile_off + next * S, S * 4);
                    bulk_prefetch_l2(p.tile_cnt + next * S, S * 2);
                }
            }
            // ---- (1) bins: thread t owns 8 consecutive buckets per round ----------------------
            {
                const uint32_t chunk = tile / p.tiles_per_chunk;
                const uint16_t *cnt = p.tile_cnt + (size_t) tile * S;
                const uint32_t *toff = p.tile_off + (size_t) tile * S, *crow = p.rows + (size_t) chunk * S;
                uint32_t carry = 0;
                for (uint32_t base = 0; base < S; base += THREADS * 8) {
                    const uint32_t b0 = base + tid * 8;
                    uint32_t c[8], g[8];
                    #pragma unroll
                    for (uint32_t j = 0; j < 8; ++j) { c[j] = 0; g[j] = 0; }
                    if (b0 < S) {
                        const uint4 cc = __ldg(reinterpret_cast<const uint4 *>(cnt + b0));
                        const uint4 t0 = __ldg(reinterpret_cast<const uint4 *>(toff + b0)),
                                    t1 = __ldg(reinterpret_cast<const uint4 *>(toff + b0 + 4)),
                                    r0 = __ldg(reinterpret_cast<const uint4 *>(crow + b0)),
                                    r1 = __ldg(reinterpret_cast<const uint4 *>(crow + b0 + 4));
                        uint32_t s8[8];
                        #pragma unroll
                        for (uint32_t j = 0; j < 8; ++j) s8[j] = b0 + j < p.bucket_count ? __ldg(p.bucket_start + b0 + j) : 0u;
                        c[0] = cc.x & 0xffffu; c[1] = cc.x >> 16; c[2] = cc.y & 0xffffu; c[3] = cc.y >> 16;
                        c[4] = cc.z & 0xffffu; c[5] = cc.z >> 16; c[6] = cc.w & 0xffffu; c[7] = cc.w >> 16;
                        g[0] = t0.x + r0.x + s8[0]; g[1] = t0.y + r0.y + s8[1]; g[2] = t0.z + r0.z + s8[2]; g[3] = t0.w + r0.w + s8[3];
                        g[4] = t1.x + r1.x + s8[4]; g[5] = t1.y + r1.y + s8[5]; g[6] = t1.z + r1.z + s8[6]; g[7] = t1.w + r1.w + s8[7];
                    }
                    uint32_t sum = 0;
                    #pragma unroll
                    for (uint32_t j = 0; j < 8; ++j) sum += c[j];
                    uint32_t incl = sum;
                    #pragma unroll
                    for (uint32_t d = 1; d < 32; d <<= 1) {
                        const uint32_t t = shfl_up(incl, d);
                        if (lane >= d) incl += t;
                    }
                    if (lane == 31) warp_sum[warp] = incl;
                    __syncthreads();
                    uint32_t wbase = 0, total = 0;
                    #pragma unroll
                    for (uint32_t w = 0; w < WARPS; ++w) {
                        if (w == warp) wbase = total;
                        total += warp_sum[w];
                    }
                    uint32_t run = carry + wbase + incl - sum;
                    carry += total;
                    if (b0 < S) {
                        uint32_t st[8];
                        #pragma unroll
                        for (uint32_t j = 0; j < 8; ++j) { st[j] = run; g[j] -= run; run += c[j]; }
                        *reinterpret_cast<uint4 *>(cursor + b0) = make_uint4(st[0], st[1], st[2], st[3]);
                        *reinterpret_cast<uint4 *>(cursor + b0 + 4) = make_uint4(st[4], st[5], st[6], st[7]);
                        *reinterpret_cast<uint4 *>(delta + b0) = make_uint4(g[0], g[1], g[2], g[3]);
                        *reinterpret_cast<uint4 *>(delta + b0 + 4) = make_uint4(g[4], g[5], g[6], g[7]);
                    }
                    __syncthreads();
                }
            }
            // ---- (2) keys: slot from the bucket's cursor, entry stored at the slot ---------------
            #pragma unroll
            for (uint32_t k = 0; k < KPT; ++k) {
                const uint32_t local = vec ? ((k / 4) * THREADS + tid) * 4 + (k & 3u) : k * THREADS + tid;
                if constexpr (PACKED) {
                    const uint32_t b = (key[k >> 1] >> (16 * (k & 1u))) & 0xffffu;
                    if (b != 0xffffu)
                        sorted[atomicAdd(cursor + b, 1u)] = (b << 16) | local;
                } else {
                    if (key[k] != 0xffffffffu)
                        sorted[atomicAdd(cursor + key[k], 1u)] = (key[k] << 16) | local;
                }
            }
        } else {
            for (uint32_t b = tid; b < S; b += THREADS) cursor[b] = 0;       // empty tile: every run ends at 0
        }
        cluster.sync();             // both tiles are ranked; their shared memory is readable cluster-wide

        // ---- (3) my half of the buckets, both tiles: one warp per bucket, 32 buckets per round -----
        const uint32_t idx0[2] = { p.index_base + (uint32_t) ((uint64_t) (2 * pair) * TILE),
                                   p.index_base + (uint32_t) ((uint64_t) (2 * pair + 1) * TILE) };
        for (uint32_t cb = h0 + warp * 32; cb < h1; cb += WARPS * 32) {
            const uint32_t b = cb + lane;
            const bool valid = b < h1;
            // run descriptors of bucket b in both tiles: start | count << 16, and the delta
            uint32_t desc[2], dl[2];
            #pragma unroll
            for (uint32_t q = 0; q < 2; ++q) {
                const uint32_t *cur_q = tile_smem[q], *delta_q = tile_smem[q] + S;
                const uint32_t end = valid ? cur_q[b] : 0u;
                uint32_t start = shfl_up(end, 1);
                if (lane == 0) start = cb ? cur_q[cb - 1] : 0u;
                desc[q] = valid ? (start | ((end - start) << 16)) : 0u;
                dl[q] = valid ? delta_q[b] : 0u;
            }
            const uint32_t nb = min(32u, h1 - cb);
            // four buckets per step: the (remote) loads of all four are in flight before the first store
            for (uint32_t k0 = 0; k0 < nb; k0 += 4) {
                uint32_t e[4], at[4], tot[4], cAs[4], sAs[4], gAs[4], sBs[4], gBs[4];
                #pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    const uint32_t k = min(k0 + u, 31u);
                    const uint32_t dA = k0 + u < nb ? shfl_idx(desc[0], k) : 0u, dB = k0 + u < nb ? shfl_idx(desc[1], k) : 0u;
                    const uint32_t sA = dA & 0xffffu, cA = dA >> 16, sB = dB & 0xffffu, cB = dB >> 16;
                    const uint32_t gA = shfl_idx(dl[0], k) + sA, gB = shfl_idx(dl[1], k) + sB;
                    tot[u] = cA + cB; cAs[u] = cA; sAs[u] = sA; gAs[u] = gA; sBs[u] = sB; gBs[u] = gB;
                    const bool second = lane >= cA;
                    const uint32_t o = second ? lane - cA : lane;
                    e[u] = 0;
                    if (lane < tot[u])
                        e[u] = (tile_smem[second ? 1 : 0] + 2 * S)[(second ? sB : sA) + o];
                    at[u] = (second ? gB : gA) + o;
                    e[u] = idx0[second ? 1 : 0] + (e[u] & 0xffffu);
                }
                #pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    if (lane < tot[u])
                        p.perm[at[u]] = e[u];
                    // rare: more than 32 entries in the two runs together
                    for (uint32_t off = lane + 32; off < tot[u]; off += 32) {
                        const bool second = off >= cAs[u];
                        const uint32_t o = second ? off - cAs[u] : off;
                        const uint32_t v = (tile_smem[second ? 1 : 0] + 2 * S)[(second ? sBs[u] : sAs[u]) + o];
                        p.perm[(second ? gBs[u] : gAs[u]) + o] = idx0[second ? 1 : 0] + (v & 0xffffu);
                    }
                }
            }
        }
        cluster.sync();             // the other CTA has read my staged entries: the buffers are free again
    }
}

#endif // DRJIT_B200_EXPERIMENTS (pair kernel)

#if defined(DRJIT_B200_EXPERIMENTS)
// ---------------------------------------------------------------------------
//  Unordered tile scatter with 16-bit staging entries (EXPERIMENTAL: DRJIT_B200_MKPERM_KPT=60)
// ---------------------------------------------------------------------------
//  DESIGN.md section 8.1, step (1). The staged entry is the key's 16-bit local index only, which
//  lets a tile hold 60 Ki keys (runs per bucket and tile are what the store path of the copy-out is
//  bound by). The copy-out recovers a slot's bucket with a rank query instead of reading it from
//  the entry: one bit per non-empty bucket start in a bitmap over the tile's slots, a 16-bit prefix
//  count per bitmap word, and a table rank -> bucket. A warp copies 32 aligned slots, so the bitmap
//  word and the prefix count are warp-uniform loads and the rank is one popcount.
//  Keys are not held in registers: they are ranked in groups of 20 per thread (5 x LDG.128).
//  Compiled only with -DDRJIT_B200_EXPERIMENTS and selected only by the environment variable above
//  (scripts/gpu_mkperm_kpt.sh is its acceptance run); not part of the shipped library.
template <uint32_t THREADS, uint32_t KPT>
__global__ void __launch_bounds__(THREADS, 1)
mkperm_tile_scatter16_kernel(const MkpermTileParams p) {
    constexpr uint32_t TILE = THREADS * KPT, WARPS = THREADS / 32, WORDS = TILE / 32, GROUP = 5;
    static_assert(TILE < 65536, "tile counts and local indices are 16-bit values");
    static_assert(KPT % (4 * GROUP) == 0, "keys are ranked in groups of 20 per thread");
    static_assert(WORDS <= 2 * THREADS, "two bitmap words per thread in the prefix count");
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t S = p.stride;
    uint32_t *cursor = smem;                                    // [S] next free slot of the bucket
    uint32_t *delta = smem + S;                                 // [S] final position of the run minus its local start
    uint32_t *bitmap = smem + 2 * S;                            // [WORDS] bit s: a non-empty bucket starts at slot s
    uint16_t *wrank = reinterpret_cast<uint16_t *>(bitmap + WORDS);       // [WORDS] set bits before the word
    uint16_t *segb = wrank + WORDS;                             // [S] rank -> bucket
    uint16_t *sorted = segb + S;                                // [TILE] local index, ordered by bucket
    __shared__ uint32_t warp_sum[WARPS];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u, last = p.bucket_count - 1;

    for (uint32_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const uint64_t tile_base = (uint64_t) tile * TILE;
        const uint32_t n_tile = (uint32_t) min((uint64_t) TILE, (uint64_t) p.size - tile_base);
        const bool vec = p.vec && n_tile == TILE;
        const uint4 *kv = reinterpret_cast<const uint4 *>(p.values + tile_base);

        // first group of keys: in flight during the bin phase
        Vec16<uint32_t> kq[GROUP];
        if (vec) {
            #pragma unroll
            for (uint32_t i = 0; i < GROUP; ++i) kq[i] = ld_stream<uint32_t>(kv + i * THREADS + tid);
        }
        if (tid == 0) {
            const uint64_t next = (uint64_t) tile + gridDim.x;
            if (next < p.tiles) {
                if (p.vec && (next + 1) * TILE <= p.size)
                    bulk_prefetch_l2(p.values + next * TILE, TILE * 4);
                bulk_prefetch_l2(p.tile_off + next * S, S * 4);
                bulk_prefetch_l2(p.tile_cnt + next * S, S * 2);
            }
        }
        for (uint32_t w = tid; w < WORDS; w += THREADS) bitmap[w] = 0;
        __syncthreads();

        // ---- (1) bins: thread t owns 8 consecutive buckets per round; boundary bits ----------
        {
            const uint32_t chunk = tile / p.tiles_per_chunk;
            const uint16_t *cnt = p.tile_cnt + (size_t) tile * S;
            const uint32_t *toff = p.tile_off + (size_t) tile * S, *crow = p.rows + (size_t) chunk * S;
            uint32_t carry = 0;
            for (uint32_t base = 0; base < S; base += THREADS * 8) {
                const uint32_t b0 = base + tid * 8;
                uint32_t c[8], g[8];
                #pragma unroll
                for (uint32_t j = 0; j < 8; ++j) { c[j] = 0; g[j] = 0; }
                if (b0 < S) {
                    const uint4 cc = __ldg(reinterpret_cast<const uint4 *>(cnt + b0));
                    const uint4 t0 = __ldg(reinterpret_cast<const uint4 *>(toff + b0)),
                                t1 = __ldg(reinterpret_cast<const uint4 *>(toff + b0 + 4)),
                                r0 = __ldg(reinterpret_cast<const uint4 *>(crow + b0)),
                                r1 = __ldg(reinterpret_cast<const uint4 *>(crow + b0 + 4));
                    uint32_t s8[8];
                    #pragma unroll
                    for (uint32_t j = 0; j < 8; ++j) s8[j] = b0 + j < p.bucket_count ? __ldg(p.bucket_start + b0 + j) : 0u;
                    c[0] = cc.x & 0xffffu; c[1] = cc.x >> 16; c[2] = cc.y & 0xffffu; c[3] = cc.y >> 16;
                    c[4] = cc.z & 0xffffu; c[5] = cc.z >> 16; c[6] = cc.w & 0xffffu; c[7] = cc.w >> 16;
                    g[0] = t0.x + r0.x + s8[0]; g[1] = t0.y + r0.y + s8[1]; g[2] = t0.z + r0.z + s8[2]; g[3] = t0.w + r0.w + s8[3];
                    g[4] = t1.x + r1.x + s8[4]; g[5] = t1.y + r1.y + s8[5]; g[6] = t1.z + r1.z + s8[6]; g[7] = t1.w + r1.w + s8[7];
                }
                uint32_t sum = 0;
                #pragma unroll
                for (uint32_t j = 0; j < 8; ++j) sum += c[j];
                uint32_t incl = sum;
                #pragma unroll
                for (uint32_t d = 1; d < 32; d <<= 1) {
                    const uint32_t t = shfl_up(incl, d);
                    if (lane >= d) incl += t;
                }
                if (lane == 31) warp_sum[warp] = incl;
                __syncthreads();
                uint32_t wbase = 0, total = 0;
                #pragma unroll
                for (uint32_t w = 0; w < WARPS; ++w) {
                    if (w == warp) wbase = total;
                    total += warp_sum[w];
                }
                uint32_t run = carry + wbase + incl - sum;
                carry += total;
                if (b0 < S) {
                    uint32_t st[8];
                    #pragma unroll
                    for (uint32_t j = 0; j < 8; ++j) {
                        st[j] = run; g[j] -= run;
                        if (c[j]) atomicOr(bitmap + (run >> 5), 1u << (run & 31u));
                        run += c[j];
                    }
                    *reinterpret_cast<uint4 *>(cursor + b0) = make_uint4(st[0], st[1], st[2], st[3]);
                    *reinterpret_cast<uint4 *>(cursor + b0 + 4) = make_uint4(st[4], st[5], st[6], st[7]);
                    *reinterpret_cast<uint4 *>(delta + b0) = make_uint4(g[0], g[1], g[2], g[3]);
                    *reinterpret_cast<uint4 *>(delta + b0 + 4) = make_uint4(g[4], g[5], g[6], g[7]);
                }
                __syncthreads();
            }
        }

        // ---- (2) set bits before every bitmap word (two words per thread) ---------------------
        {
            const uint32_t w0 = 2 * tid, w1 = w0 + 1;
            const uint32_t p0 = w0 < WORDS ? (uint32_t) __popc(bitmap[w0]) : 0u,
                           p1 = w1 < WORDS ? (uint32_t) __popc(bitmap[w1]) : 0u;
            uint32_t incl = p0 + p1;
            #pragma unroll
            for (uint32_t d = 1; d < 32; d <<= 1) {
                const uint32_t t = shfl_up(incl, d);
                if (lane >= d) incl += t;
            }
            if (lane == 31) warp_sum[warp] = incl;
            __syncthreads();
            uint32_t wbase = 0;
            #pragma unroll
            for (uint32_t w = 0; w < WARPS; ++w)
                if (w < warp) wbase += warp_sum[w];
            const uint32_t excl = wbase + incl - p0 - p1;
            if (w0 < WORDS) wrank[w0] = (uint16_t) excl;
            if (w1 < WORDS) wrank[w1] = (uint16_t) (excl + p0);
        }
        __syncthreads();

        // ---- (3) rank -> bucket, from the untouched cursors ------------------------------------
        for (uint32_t b = tid; b < S; b += THREADS) {
            const uint32_t st = cursor[b], en = b + 1 < S ? cursor[b + 1] : n_tile;
            if (en != st) {
                const uint32_t r = wrank[st >> 5] + (uint32_t) __popc(bitmap[st >> 5] & ((1u << (st & 31u)) - 1u));
                segb[r] = (uint16_t) b;
            }
        }
        __syncthreads();

        // ---- (4) ranking: slot from the bucket's cursor, local index stored at the slot ---------
        if (vec) {
            #pragma unroll
            for (uint32_t grp = 0; grp < KPT / (4 * GROUP); ++grp) {
                Vec16<uint32_t> nx[GROUP];
                if (grp + 1 < KPT / (4 * GROUP)) {
                    #pragma unroll
                    for (uint32_t i = 0; i < GROUP; ++i)
                        nx[i] = ld_stream<uint32_t>(kv + ((grp + 1) * GROUP + i) * THREADS + tid);
                }
                #pragma unroll
                for (uint32_t i = 0; i < GROUP; ++i) {
                    const uint32_t local = ((grp * GROUP + i) * THREADS + tid) * 4;
                    #pragma unroll
                    for (uint32_t e = 0; e < 4; ++e)
                        sorted[atomicAdd(cursor + min(kq[i].v[e], last), 1u)] = (uint16_t) (local + e);
                }
                #pragma unroll
                for (uint32_t i = 0; i < GROUP; ++i) kq[i] = nx[i];
            }
        } else {
            for (uint32_t i = tid; i < n_tile; i += THREADS)
                sorted[atomicAdd(cursor + min(__ldg(p.values + tile_base + i), last), 1u)] = (uint16_t) i;
        }
        __syncthreads();

        // ---- (5) copy-out: the bucket of slot j is segb[number of bucket starts <= j, minus 1] ---
        const uint32_t idx0 = p.index_base + (uint32_t) tile_base;
        #pragma unroll 4
        for (uint32_t j = tid; j < n_tile; j += THREADS) {
            const uint32_t w = j >> 5;
            const uint32_t r = wrank[w] + (uint32_t) __popc(bitmap[w] & lanemask_le()) - 1u;
            p.perm[delta[segb[r]] + j] = idx0 + sorted[j];
        }
        __syncthreads();
    }
}

#endif // DRJIT_B200_EXPERIMENTS

// ---------------------------------------------------------------------------
//  Stable tile scatter (bucket counts for which the reference guarantees a stable permutation)
// ---------------------------------------------------------------------------
//  jit.h:2404-2406: the reference's "tiny" variant -- bucket_count * 4 bytes * 32 warps fit into
//  shared memory, i.e. up to 1816 buckets on this GPU -- produces a *stable* permutation, and
//  dr.sort / dr.argsort rely on it: they are LSD radix sorts made of block_mkperm passes with 256
//  buckets (drjit/__init__.py:1698-1772). For these bucket counts the tile path therefore ranks
//  keys in input order: every warp owns a contiguous segment of the tile and walks it 32
//  consecutive keys at a time; lanes holding the same key are found with one ballot per key bit
//  (2 instructions per bit instead of a ~64-cycle match.any), the rank inside the step is a
//  popcount, and the running position of (warp, bucket) lives in a warp-private counter row.
//  Order inside a bucket = tile order (tile_off) > warp order (column prefix over the per-warp
//  histograms) > step order > lane order = input order. Everything else (K1, K2, copy-out, L2
//  prefetch) is shared with the unordered kernel above.
/// peers &= (lanes whose key agrees with mine in the bit `mask`): test, ballot, select, one LOP3
__device__ __forceinline__ uint32_t match_bit(uint32_t peers, uint32_t key, uint32_t mask) {
    const bool one = key & mask;
    const uint32_t v = __ballot_sync(kFullMask, one), kb = one ? 0xffffffffu : 0u;
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x90;" : "=r"(d) : "r"(peers), "r"(v), "r"(kb));    // peers & ~(v ^ kb)
    return d;
}

template <uint32_t THREADS, uint32_t KEY_BITS>
__global__ void __launch_bounds__(THREADS, 1)
mkperm_tile_scatter_stable_kernel(const MkpermTileParams p) {
    constexpr uint32_t TILE = THREADS * kTileKeysPerThread, WARPS = THREADS / 32, SEG = TILE / WARPS;
    static_assert(TILE <= 65536, "local indices are packed into 16 bits");
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t S = p.stride;
    uint32_t *whist = smem;                    // [WARPS][S] per-warp counts -> running tile-local positions
    uint32_t *delta = smem + WARPS * S;        // [S] final position of the bucket's run minus its local start
    uint32_t *sorted = delta + S;              // [TILE] (bucket << 16 | local index), ordered by bucket
    __shared__ uint32_t warp_sum[WARPS];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u, last = p.bucket_count - 1;
    uint32_t *mine = whist + warp * S;

    for (uint32_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const uint64_t tile_base = (uint64_t) tile * TILE;
        const uint32_t n_tile = (uint32_t) min((uint64_t) TILE, (uint64_t) p.size - tile_base);

        // keys of step s of this warp's segment: seg + s * 32 + lane (coalesced, input order)
        uint32_t key[kTileKeysPerThread];
        #pragma unroll
        for (uint32_t s = 0; s < kTileKeysPerThread; ++s) {
            const uint32_t i = warp * SEG + s * 32 + lane;
            key[s] = i < n_tile ? min(__ldg(p.values + tile_base + i), last) : 0xffffffffu;
        }
        if (tid == 0) {
            const uint64_t next = (uint64_t) tile + gridDim.x;
            if (p.n_pay && tile == blockIdx.x) tile_prefetch_payloads(p, tile, TILE);
            if (next < p.tiles) {
                if (p.vec && (next + 1) * TILE <= p.size)
                    bulk_prefetch_l2(p.values + next * TILE, TILE * 4);
                bulk_prefetch_l2(p.tile_off + next * S, S * 4);
                if (p.n_pay) tile_prefetch_payloads(p, next, TILE);
            }
        }

        // ---- (1) per-warp histograms of the segments -------------------------------------------
        for (uint32_t b = lane; b < S; b += 32) mine[b] = 0;
        __syncwarp();
        #pragma unroll
        for (uint32_t s = 0; s < kTileKeysPerThread; ++s)
            if (key[s] != 0xffffffffu) atomicAdd(mine + key[s], 1u);
        __syncthreads();

        // ---- (2) bins: prefix over the warps of each bucket, then over the buckets ----------------
        {
            const uint32_t chunk = tile / p.tiles_per_chunk;
            const uint32_t *toff = p.tile_off + (size_t) tile * S, *crow = p.rows + (size_t) chunk * S;
            uint32_t carry = 0;
            for (uint32_t base = 0; base < S; base += THREADS) {
                const uint32_t b = base + tid;
                uint32_t tot = 0, goff = 0;
                if (b < S) {
                    goff = __ldg(toff + b) + __ldg(crow + b) + (b < p.bucket_count ? __ldg(p.bucket_start + b) : 0u);
                    #pragma unroll 8
                    for (uint32_t w = 0; w < WARPS; ++w) {
                        const uint32_t c = whist[w * S + b];
                        whist[w * S + b] = tot;
                        tot += c;
                    }
                }
                uint32_t incl = tot;
                #pragma unroll
                for (uint32_t d = 1; d < 32; d <<= 1) {
                    const uint32_t t = shfl_up(incl, d);
                    if (lane >= d) incl += t;
                }
                if (lane == 31) warp_sum[warp] = incl;
                __syncthreads();
                uint32_t wbase = 0, total = 0;
                #pragma unroll
                for (uint32_t w = 0; w < WARPS; ++w) {
                    if (w == warp) wbase = total;
                    total += warp_sum[w];
                }
                const uint32_t start = carry + wbase + incl - tot;      // tile-local start of bucket b
                carry += total;
                if (b < S) {
                    #pragma unroll 8
                    for (uint32_t w = 0; w < WARPS; ++w) whist[w * S + b] += start;
                    delta[b] = goff - start;
                }
                __syncthreads();
            }
        }

        // ---- (3) ranking walk in input order -------------------------------------------------------
        // (FULL: every lane holds a key; only the last tile of the array can be ragged)
        auto walk = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            #pragma unroll
            for (uint32_t s = 0; s < kTileKeysPerThread; ++s) {
                const uint32_t k = key[s];
                const bool valid = FULL || k != 0xffffffffu;
                uint32_t peers = FULL ? kFullMask : __ballot_sync(kFullMask, valid);
                #pragma unroll
                for (uint32_t bit = 0; bit < KEY_BITS; ++bit)
                    peers = match_bit(peers, k, 1u << bit);
                const uint32_t rank = __popc(peers & lanemask_lt());
                uint32_t pos = 0;
                if (valid) pos = mine[k] + rank;
                __syncwarp();
                if (valid && rank == 0) mine[k] = pos + __popc(peers);  // lowest lane of the group
                __syncwarp();
                if (valid) sorted[pos] = (k << 16) | (warp * SEG + s * 32 + lane);
            }
        };
        if (n_tile == TILE) walk(std::true_type{});
        else                walk(std::false_type{});
        __syncthreads();

        // ---- (4) runs of equal buckets are contiguous in `sorted` and in `perm` --------------------
        const uint32_t idx0 = p.index_base + (uint32_t) tile_base;
        if (p.n_pay) {
            tile_copy_out_payloads<THREADS>(p, sorted, delta, tile_base, n_tile, idx0);
        } else {
            #pragma unroll 4
            for (uint32_t j = tid; j < n_tile; j += THREADS) {
                const uint32_t e = sorted[j];
                p.perm[delta[e >> 16] + j] = idx0 + (e & 0xffffu);
            }
        }
        __syncthreads();
    }
}

/// Payload gather for the paths without a fused copy-out (small inputs, several sorting groups)
__global__ void mkperm_gather_kernel(const uint32_t *perm, uint32_t size, uint32_t index_base, const uint32_t *in, uint32_t *out) {
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (uint64_t) gridDim.x * blockDim.x)
        out[i] = __ldg(in + (perm[i] - index_base));
}

// ---------------------------------------------------------------------------
//  Host side
// ---------------------------------------------------------------------------
static MkpermMode pick_mode(uint32_t bucket_count, uint32_t smem_budget, uint32_t &warps) {
    const uint64_t bytes = (uint64_t) bucket_count * 4;
    MkpermMode mode;
    uint32_t w = (uint32_t) std::min<uint64_t>(32, smem_budget / bytes);
    if (w >= 8) { mode = MkpermMode::Warp; warps = w; }
    else if (bytes <= smem_budget) { mode = MkpermMode::Cta; warps = 32; }
    else { mode = MkpermMode::Global; warps = 32; }

#if defined(DRJIT_B200_EXPERIMENTS)
    // A/B measurements: DRJIT_B200_MKPERM_MODE=warp|cta|global
    if (const char *env = getenv("DRJIT_B200_MKPERM_MODE")) {
        if (!strcmp(env, "cta") && bytes <= smem_budget) { mode = MkpermMode::Cta; warps = 32; }
        else if (!strcmp(env, "global")) { mode = MkpermMode::Global; warps = 32; }
        else if (!strcmp(env, "warp") && w >= 1) { mode = MkpermMode::Warp; warps = w; }
    }
#endif
    return mode;
}

template <MkpermMode Mode>
static void launch_phases(cudaStream_t stream, MkpermParams &p, uint32_t threads, uint32_t smem,
                          uint32_t *offsets_dev, uint32_t *unique_dev, uint32_t *hist_out,
                          cudaEvent_t table_ready, const MkpermPeer *peer, bool sort_table) {
    const uint32_t grid = p.ctas_per_group * p.n_groups;
    if (smem > 48 * 1024) {
        DJB_CUDA_CHECK(cudaFuncSetAttribute(mkperm_histogram_kernel<Mode>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        DJB_CUDA_CHECK(cudaFuncSetAttribute(mkperm_scatter_kernel<Mode>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    }
    mkperm_histogram_kernel<Mode><<<grid, threads, smem, stream>>>(p);
    DJB_POST_LAUNCH();
    if (Mode != MkpermMode::Global) {
        const uint32_t tiles = ceil_div(p.bucket_count, 32);
        mkperm_column_scan_kernel<<<tiles * p.n_groups, 256, 0, stream>>>(p, tiles);
        DJB_POST_LAUNCH();
    }
    launch_bucket_scan(stream, p, offsets_dev, unique_dev, hist_out, peer, sort_table);
    if (table_ready)
        DJB_CUDA_CHECK(cudaEventRecord(table_ready, stream)); // cuda_ts.cpp:953 (before phase 4)
    mkperm_scatter_kernel<Mode><<<grid, threads, smem, stream>>>(p);
    DJB_POST_LAUNCH();
}



#if defined(DRJIT_B200_EXPERIMENTS)
/// Integer environment variable, read once (experiments builds only)
static int env_int(const char *name, int fallback) {
    const char *env = getenv(name);
    return env ? atoi(env) : fallback;
}
#endif

static bool use_tile_path(uint32_t n_groups, uint32_t size, uint32_t bucket_count) {
#if defined(DRJIT_B200_EXPERIMENTS)
    static const int enabled = env_int("DRJIT_B200_MKPERM_TILES", 1);  // =0 disables the tile path
#else
    constexpr int enabled = 1;
#endif
    const uint64_t tiles = ceil_div64(size, 512 * kTileKeysPerThread);
    return enabled && n_groups == 1 && bucket_count <= kTileMaxBuckets && size >= (1u << 18) &&
           tiles * bucket_count * 6 <= ((uint64_t) 2 << 30);
}

#if defined(DRJIT_B200_EXPERIMENTS)
/// Cluster-of-two scatter kernel (mkperm_tile_scatter_pair_kernel): measured slower, kept for A/B
/// runs in experiments builds (DRJIT_B200_MKPERM_PAIR=1)
static bool use_pair_kernel(const MkpermTileParams &t) {
    static const int pair = env_int("DRJIT_B200_MKPERM_PAIR", 0);
    return pair != 0 && t.n_pay == 0;
}
#endif

template <uint32_t THREADS, uint32_t KEY_BITS>
static void launch_stable_scatter(cudaStream_t stream, const MkpermTileParams &t, uint32_t grid, uint32_t smem, uint32_t smem_max) {
    static std::atomic<bool> configured_on[kMaxDevices] = {};       // (function attributes are per device)
    std::atomic<bool> &configured = configured_on[device_props().device % kMaxDevices];
    if (!configured.load(std::memory_order_acquire)) {     // (idempotent: a race only repeats the call)
        DJB_CUDA_CHECK(cudaFuncSetAttribute(mkperm_tile_scatter_stable_kernel<THREADS, KEY_BITS>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_max));
        configured.store(true, std::memory_order_release);
    }
    mkperm_tile_scatter_stable_kernel<THREADS, KEY_BITS><<<grid, THREADS, smem, stream>>>(t);
}

template <uint32_t THREADS, bool STABLE, uint32_t KPT = kTileKeysPerThread, bool PAYS = false>
static uint32_t mkperm_tiles(cudaStream_t stream, const uint32_t *values, uint32_t size,
                             uint32_t bucket_count, uint32_t index_base, uint32_t *perm,
                             uint32_t *offsets, uint32_t *hist_out, const MkpermPeer *peer, const MkpermExtras &ex) {
    static_assert(!STABLE || KPT == kTileKeysPerThread, "the stable kernel has 32 keys per thread");
    constexpr uint32_t TILE = THREADS * KPT;
    const DeviceProps &dev = device_props();
    MkpermTileParams t{};
    t.values = values; t.perm = perm; t.size = size; t.bucket_count = bucket_count;
    t.stride = (bucket_count + 7) / 8 * 8;              // rows are read with 128-bit loads
    t.index_base = index_base;
    t.tiles = (uint32_t) ceil_div64(size, TILE);
    t.vec = ((uintptr_t) values % 16) == 0;
    t.n_pay = ex.n_pay;
    for (uint32_t k = 0; k < ex.n_pay; ++k) { t.pay_in[k] = ex.pay_in[k]; t.pay_out[k] = ex.pay_out[k]; }
#if defined(DRJIT_B200_EXPERIMENTS)
    {
        static const int debug = env_int("DRJIT_B200_MKPERM_DEBUG", 0);
        t.debug = (uint8_t) debug;
    }
#endif

#if defined(DRJIT_B200_EXPERIMENTS)
    constexpr bool STAGE16 = !STABLE && KPT >= 60;       // 16-bit staging entries (mkperm_tile_scatter16_kernel)
#else
    constexpr bool STAGE16 = false;
#endif
    const uint32_t hist_smem = t.stride * 8,
                   scatter_smem = STABLE ? (THREADS / 32 + 1) * t.stride * 4 + TILE * 4
                                : STAGE16 ? t.stride * 10 + TILE / 32 * 6 + TILE * 2
                                          : t.stride * 8 + TILE * 4 * (PAYS ? 2 : 1);
    uint32_t chunks = std::min(t.tiles, dev.sm_count * 2);
    t.tiles_per_chunk = ceil_div(t.tiles, chunks);
    chunks = ceil_div(t.tiles, t.tiles_per_chunk);

    // the column / bucket scan kernels work on MkpermParams: one group, `chunks` rows
    MkpermParams p{};
    p.values = values; p.perm = perm; p.size = size; p.block_size = size;
    p.bucket_count = bucket_count; p.n_groups = 1; p.rows_per_group = chunks;
    p.row_stride = t.stride;

    Scratch scratch(stream);
    const size_t off_bytes = (size_t) t.tiles * t.stride * 4,
                 cnt_bytes = (size_t) t.tiles * t.stride * 2,
                 rows_bytes = (size_t) chunks * t.stride * 4,
                 totals_bytes = (size_t) bucket_count * 4;
    auto r256 = [](size_t v) { return (v + 255) & ~(size_t) 255; };
    scratch.reserve(r256(off_bytes) + r256(cnt_bytes) + r256(rows_bytes) + r256(totals_bytes) + 512);
    t.tile_off = (uint32_t *) scratch.device(off_bytes);
    t.tile_cnt = (uint16_t *) scratch.device(cnt_bytes);
    t.rows = p.rows = (uint32_t *) scratch.device(rows_bytes);
    p.totals = (uint32_t *) scratch.device(totals_bytes);
    t.bucket_start = p.totals;

    const bool want_table = offsets != nullptr;
    uint32_t *offsets_dev = nullptr, *unique_dev = nullptr;
    uint32_t *pinned = scratch.pinned_words();
    if (want_table) {
        cudaError_t rv = cudaHostGetDevicePointer((void **) &offsets_dev, offsets, 0);
        if (rv != cudaSuccess) {
            (void) cudaGetLastError();
            raise(DRJIT_B200_EINVAL, "jit_block_mkperm(): 'offsets' must point to host-pinned "
                                     "(device-mapped) memory!");
        }
        DJB_CUDA_CHECK(cudaHostGetDevicePointer((void **) &unique_dev, pinned + 1, 0));
    }

    static std::atomic<bool> configured_on[kMaxDevices] = {};       // (function attributes are per device)
    std::atomic<bool> &configured = configured_on[dev.device % kMaxDevices];
    if (!configured.load(std::memory_order_acquire)) {
        DJB_CUDA_CHECK(cudaFuncSetAttribute(mkperm_tile_hist_kernel<THREADS, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int) (kTileMaxBuckets * 8)));
#if defined(DRJIT_B200_EXPERIMENTS)
        if constexpr (STAGE16)
            DJB_CUDA_CHECK(cudaFuncSetAttribute(mkperm_tile_scatter16_kernel<THREADS, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int) (dev.smem_optin - 1024)));
        else
#endif
        if (!STABLE)
            DJB_CUDA_CHECK(cudaFuncSetAttribute(mkperm_tile_scatter_kernel<THREADS, KPT, PAYS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int) std::min<uint32_t>(kTileMaxBuckets * 8 + TILE * 4 * (PAYS ? 2 : 1), dev.smem_optin - 1024)));
        configured.store(true, std::memory_order_release);
    }
    if (!STABLE && scatter_smem > dev.smem_optin - 1024)
        raise(DRJIT_B200_EFATAL, "jit_block_mkperm(): internal error (tile does not fit into shared memory)");

    mkperm_tile_hist_kernel<THREADS, KPT><<<chunks, THREADS, hist_smem, stream>>>(t);
    DJB_POST_LAUNCH();
    const uint32_t col_tiles = ceil_div(bucket_count, 32);
    mkperm_column_scan_kernel<<<col_tiles, 256, 0, stream>>>(p, col_tiles);
    DJB_POST_LAUNCH();
    launch_bucket_scan(stream, p, offsets_dev, unique_dev, hist_out, peer, ex.sort_table);
    cudaEvent_t ev = want_table ? thread_event() : nullptr;
    if (ev)
        DJB_CUDA_CHECK(cudaEventRecord(ev, stream));       // cuda_ts.cpp:953 (before the scatter pass)
    if (STABLE) {
        // one ballot per key bit, unrolled: instantiated for 4 / 8 / 9 (32 Ki-key tiles) and 11 bits
        const uint32_t grid = std::min(t.tiles, dev.sm_count), smem_max = dev.smem_optin - 1024;  // (static part: warp_sum)
        if (THREADS == 512)          launch_stable_scatter<512, 11>(stream, t, grid, scatter_smem, smem_max);
        else if (bucket_count <= 16)  launch_stable_scatter<1024, 4>(stream, t, grid, scatter_smem, smem_max);
        else if (bucket_count <= 256) launch_stable_scatter<1024, 8>(stream, t, grid, scatter_smem, smem_max);
        else                          launch_stable_scatter<1024, 9>(stream, t, grid, scatter_smem, smem_max);
    } else {
        const uint32_t ctas = std::max(1u, std::min(1024 / THREADS, (dev.smem_optin + 1024) / (scatter_smem + 1024)));
        const uint32_t grid = std::min(t.tiles, dev.sm_count * ctas);
#if defined(DRJIT_B200_EXPERIMENTS)
        if constexpr (!STABLE && THREADS == 1024 && KPT >= 40) {
            if (use_pair_kernel(t)) {
                // clusters of two CTAs, two consecutive tiles per cluster and round
                static std::atomic<bool> pair_configured_on[kMaxDevices] = {};
                std::atomic<bool> &pc = pair_configured_on[dev.device % kMaxDevices];
                if (!pc.load(std::memory_order_acquire)) {
                    DJB_CUDA_CHECK(cudaFuncSetAttribute(mkperm_tile_scatter_pair_kernel<THREADS, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                        (int) std::min<uint32_t>(kTileMaxBuckets * 8 + TILE * 4, dev.smem_optin - 1024)));
                    pc.store(true, std::memory_order_release);
                }
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(2 * std::max(1u, std::min((t.tiles + 1) / 2, dev.sm_count / 2)));
                cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = scatter_smem; cfg.stream = stream;
                cudaLaunchAttribute attr{};
                attr.id = cudaLaunchAttributeClusterDimension;
                attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
                cfg.attrs = &attr; cfg.numAttrs = 1;
                DJB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mkperm_tile_scatter_pair_kernel<THREADS, KPT>, t));
                DJB_POST_LAUNCH();
                if (!want_table)
                    return 0;
                scratch.unlock();
                DJB_CUDA_CHECK(cudaEventSynchronize(ev));
                return pinned[1];
            }
        }
#endif
#if defined(DRJIT_B200_EXPERIMENTS)
        if constexpr (STAGE16)
            mkperm_tile_scatter16_kernel<THREADS, KPT><<<std::min(t.tiles, dev.sm_count), THREADS, scatter_smem, stream>>>(t);
        else
#endif
            mkperm_tile_scatter_kernel<THREADS, KPT, PAYS><<<grid, THREADS, scatter_smem, stream>>>(t);
    }
    DJB_POST_LAUNCH();

    if (!want_table)
        return 0;
    scratch.unlock();                                      // (never block on the GPU with the stream's lock held)
    DJB_CUDA_CHECK(cudaEventSynchronize(ev));              // cuda_ts.cpp:964-967
    return pinned[1];
}

/// An empty trailing shard of a sharded call still takes part in the histogram exchange
static uint32_t mkperm_empty_shard(cudaStream_t stream, uint32_t bucket_count, uint32_t *offsets,
                                   uint32_t *hist_out, const MkpermPeer *peer) {
    MkpermParams p{};
    p.bucket_count = bucket_count; p.n_groups = 1;
    Scratch scratch(stream);
    p.totals = (uint32_t *) scratch.device((size_t) bucket_count * 4);
    DJB_CUDA_CHECK(cudaMemsetAsync(p.totals, 0, (size_t) bucket_count * 4, stream));
    uint32_t *offsets_dev = nullptr, *unique_dev = nullptr, *pinned = scratch.pinned_words();
    if (offsets) {
        DJB_CUDA_CHECK(cudaHostGetDevicePointer((void **) &offsets_dev, offsets, 0));
        DJB_CUDA_CHECK(cudaHostGetDevicePointer((void **) &unique_dev, pinned + 1, 0));
    }
    launch_bucket_scan(stream, p, offsets_dev, unique_dev, hist_out, peer);
    if (!offsets)
        return 0;
    scratch.unlock();
    DJB_CUDA_CHECK(cudaStreamSynchronize(stream));
    return pinned[1];
}

/// Shared implementation. offsets: pinned host table (may be NULL). hist_out: optional device
/// array receiving the per-bucket counts (single group only). Returns the unique count when
/// `offsets` is given and there is a single group (after waiting for the table), else 0.
static uint32_t mkperm_impl(cudaStream_t stream, const uint32_t *values, uint32_t size,
                            uint32_t block_size, uint32_t bucket_count, uint32_t index_base,
                            uint32_t *perm, uint32_t *offsets, uint32_t *hist_out,
                            const MkpermPeer *peer = nullptr, const MkpermExtras &ex = MkpermExtras()) {
    if (bucket_count == 0) // cuda_ts.cpp:794-795 (jitc_fail)
        raise(DRJIT_B200_EFATAL, "jit_block_mkperm(): bucket_count cannot be zero!");
    if (size == 0 && peer)
        return mkperm_empty_shard(stream, bucket_count, offsets, hist_out, peer);
    if (size == 0)
        return 0;
    if (block_size == 0 || block_size > size)
        raise(DRJIT_B200_EINVAL, "jit_block_mkperm(): invalid block size (size=%u, block_size=%u)!",
              size, block_size);

    const DeviceProps &dev = device_props();
    MkpermParams p{};
    p.values = values; p.perm = perm; p.size = size; p.block_size = block_size;
    p.bucket_count = bucket_count; p.index_base = index_base;
    p.n_groups = ceil_div(size, block_size);
    while (p.key_bits < 32 && (1ull << p.key_bits) < bucket_count) ++p.key_bits;

    if (use_tile_path(p.n_groups, size, bucket_count)) {
        // Where the reference's "tiny" variant applies (bucket_count * 4 B * 32 warps fit into shared
        // memory, jit.h:2404-2406, cuda_ts.cpp:824-836) its permutation is stable and dr.sort depends
        // on that, so is ours; beyond it the reference is unordered as well and the faster unordered
        // tile kernel is used.
#if defined(DRJIT_B200_EXPERIMENTS)
        static const int force_unordered = env_int("DRJIT_B200_MKPERM_UNORDERED", 0);  // A/B measurements
#else
        constexpr int force_unordered = 0;
#endif
        const bool stable = (uint64_t) bucket_count * 4 * 32 <= dev.smem_optin && !force_unordered;
        // 32 per-warp counter rows + a 32 Ki-key tile fit up to 512 buckets; inputs that would leave
        // a quarter of the SMs without such a tile take 16 Ki-key tiles (twice as many CTAs at work:
        // 2^18..2^21 keys 45 -> 35 us, scripts/small_sizes.py)
        if (stable && bucket_count <= 512 && size >= dev.sm_count * 24576u)
            return mkperm_tiles<1024, true>(stream, values, size, bucket_count, index_base, perm, offsets, hist_out, peer, ex);
        if (stable)                             // 16 rows + a 16 Ki-key tile (up to 1816 buckets)
            return mkperm_tiles<512, true>(stream, values, size, bucket_count, index_base, perm, offsets, hist_out, peer, ex);
        // Unordered kernel: what bounds its scatter pass is the length of a bucket's run per tile
        // (profiles/r1b_microbench.txt), so the tile is as large as shared memory allows: 48 Ki keys
        // (two 16-bit keys per register) next to the two bucket rows up to 4352 buckets, 32 Ki keys
        // then 40 Ki keys up to 8192.
#if defined(DRJIT_B200_EXPERIMENTS)
        static const int kpt_env = env_int("DRJIT_B200_MKPERM_KPT", 0);   // 32|40|48|60 caps the tile (A/B measurements)
#else
        constexpr int kpt_env = 0;
#endif
        const uint32_t stride = (bucket_count + 7) / 8 * 8;
        auto fits = [&](uint32_t kpt) {
            return stride * 8 + 1024 * kpt * 4 <= dev.smem_optin - 1024 && size >= dev.sm_count * 2u * 1024u * kpt;
        };
        const uint32_t kpt = kpt_env ? (uint32_t) kpt_env : kMkpermDefaultKpt;
#if defined(DRJIT_B200_EXPERIMENTS)
        // (experimental, only by request: 60 Ki-key tiles with 16-bit staging entries)
        if (kpt_env == 60 && ex.n_pay == 0 && stride * 10 + 1024 * 60 / 32 * 6 + 1024 * 60 * 2 <= dev.smem_optin - 1024 &&
            size >= dev.sm_count * 2u * 1024u * 60u)
            return mkperm_tiles<1024, false, 60>(stream, values, size, bucket_count, index_base, perm, offsets, hist_out, peer, ex);
#endif
        // Calls that carry payload arrays (jit_var_call_reduce): half of shared memory stages the payload
        // tile, the key tile shrinks to 24 Ki (up to 4352 buckets) or 20 Ki keys (up to 8192)
        if (ex.n_pay && !kpt_env) {
            auto fits_pay = [&](uint32_t k) {
                return stride * 8 + 2 * 1024 * k * 4 <= dev.smem_optin - 1024 && size >= dev.sm_count * 2u * 1024u * k;
            };
            if (fits_pay(24))
                return mkperm_tiles<1024, false, 24, true>(stream, values, size, bucket_count, index_base, perm, offsets, hist_out, peer, ex);
            if (fits_pay(20))
                return mkperm_tiles<1024, false, 20, true>(stream, values, size, bucket_count, index_base, perm, offsets, hist_out, peer, ex);
        }
        // (two co-resident 512-thread CTAs with 20 Ki / 16 Ki-key tiles were measured and are slower:
        // 0.396 / 0.499 ms against 0.348 ms, profiles/r2o_mkperm_tile_keys.txt)
        if (kpt >= 48 && fits(48))
            return mkperm_tiles<1024, false, 48>(stream, values, size, bucket_count, index_base, perm, offsets, hist_out, peer, ex);
        if (kpt >= 40 && fits(40))
            return mkperm_tiles<1024, false, 40>(stream, values, size, bucket_count, index_base, perm, offsets, hist_out, peer, ex);
        return mkperm_tiles<1024, false>(stream, values, size, bucket_count, index_base, perm, offsets, hist_out, peer, ex);
    }

    uint32_t warps = 32;
    const uint32_t smem_budget = dev.smem_optin - 1024;
    const MkpermMode mode = pick_mode(bucket_count, smem_budget, warps);
    // Keys per row (= per warp): a row costs `bucket_count` counters to publish and scan, so it
    // should see a few times that many keys, but rows are walked 32 keys per match.any step, so
    // with few buckets shorter rows (more warps and CTAs) cut the latency of mid-sized inputs
    // (2^16 keys, 64 buckets: 81 -> ~25 us, scripts/small_sizes.py)
    const uint32_t min_row_elems = std::max(512u, std::min(2048u, 4 * bucket_count));
    // small sorting groups: do not spend more warps (= histogram rows) than the group can feed
    warps = std::max(1u, std::min(warps, ceil_div(block_size, min_row_elems)));
    const uint32_t threads = warps * 32;

    // Rows: one per warp. Spread each group over enough CTAs to fill the machine once.
    uint32_t ctas_per_group = std::max(1u, dev.sm_count / p.n_groups);
    const uint32_t max_ctas = ceil_div(block_size, min_row_elems * warps);
    ctas_per_group = std::max(1u, std::min(ctas_per_group, max_ctas));
    uint32_t row_elems = ceil_div(block_size, ctas_per_group * warps);
    row_elems = (row_elems + 127) / 128 * 128;          // keeps slices 16-byte aligned
    ctas_per_group = ceil_div(block_size, row_elems * warps);
    p.ctas_per_group = ctas_per_group;
    p.row_elems = row_elems;
    p.rows_per_group = mode == MkpermMode::Warp ? ctas_per_group * warps
                     : mode == MkpermMode::Cta ? ctas_per_group : 0;
    p.vec = ((uintptr_t) values % 16) == 0 && (block_size % 4 == 0 || p.n_groups == 1);

    Scratch scratch(stream);
    const size_t rows_bytes = (size_t) p.n_groups * p.rows_per_group * bucket_count * 4,
                 totals_bytes = (size_t) p.n_groups * bucket_count * 4;
    scratch.reserve(((rows_bytes + 255) & ~(size_t) 255) + ((totals_bytes + 255) & ~(size_t) 255) + 512);
    p.rows = (uint32_t *) scratch.device(rows_bytes);
    p.totals = (uint32_t *) scratch.device(totals_bytes);
    if (mode == MkpermMode::Global)
        DJB_CUDA_CHECK(cudaMemsetAsync(p.totals, 0, totals_bytes, stream));

    const bool want_table = offsets != nullptr && p.n_groups == 1;
    uint32_t *offsets_dev = nullptr, *unique_dev = nullptr;
    uint32_t *pinned = scratch.pinned_words();
    if (want_table) {
        // `offsets` is host-pinned memory (jit.h:2408-2411): obtain its device alias
        cudaError_t rv = cudaHostGetDevicePointer((void **) &offsets_dev, offsets, 0);
        if (rv != cudaSuccess) {
            (void) cudaGetLastError();
            raise(DRJIT_B200_EINVAL, "jit_block_mkperm(): 'offsets' must point to host-pinned "
                                     "(device-mapped) memory!");
        }
        DJB_CUDA_CHECK(cudaHostGetDevicePointer((void **) &unique_dev, pinned + 1, 0));
    }
    if (p.n_groups != 1)
        hist_out = nullptr;

    const uint32_t smem = mode == MkpermMode::Warp ? warps * bucket_count * 4
                        : mode == MkpermMode::Cta ? bucket_count * 4 : 0;
    cudaEvent_t ev = want_table ? thread_event() : nullptr;
    switch (mode) {
        case MkpermMode::Warp:
            launch_phases<MkpermMode::Warp>(stream, p, threads, smem, offsets_dev, unique_dev, hist_out, ev, peer, ex.sort_table);
            break;
        case MkpermMode::Cta:
            launch_phases<MkpermMode::Cta>(stream, p, threads, smem, offsets_dev, unique_dev, hist_out, ev, peer, ex.sort_table);
            break;
        default:
            launch_phases<MkpermMode::Global>(stream, p, threads, 0, offsets_dev, unique_dev, hist_out, ev, peer, ex.sort_table);
            break;
    }
    for (uint32_t k = 0; k < ex.n_pay; ++k) {       // (no fused copy-out on this path: one gather per payload)
        mkperm_gather_kernel<<<std::min(ceil_div(size, 1024), dev.sm_count * 8), 256, 0, stream>>>(perm, size, index_base, ex.pay_in[k], ex.pay_out[k]);
        DJB_POST_LAUNCH();
    }

    if (!want_table)
        return 0;
    scratch.unlock();
    DJB_CUDA_CHECK(cudaEventSynchronize(ev)); // cuda_ts.cpp:964-967: table valid, perm still in flight
    return pinned[1];
}

uint32_t block_mkperm(cudaStream_t stream, const uint32_t *values, uint32_t size, uint32_t block_size,
                      uint32_t bucket_count, uint32_t *perm, uint32_t *offsets) {
    return mkperm_impl(stream, values, size, block_size, bucket_count, 0, perm, offsets, nullptr);
}

/// jit_var_call_reduce (src/call.cpp:1268-1389) as one call: the permutation that groups the callable
/// IDs, the table of non-empty buckets sorted by decreasing size (the order in which the dispatcher
/// launches the callees, :1346-1356 -- sorted on the device, inside the bucket-scan kernel) and up to
/// four 32-bit argument arrays permuted on the way (pay_out[k][j] = pay_in[k][perm[j]]: written
/// where `perm` is written, which saves the separate gather pass per argument).
uint32_t call_reduce(cudaStream_t stream, const uint32_t *ids, uint32_t size, uint32_t bucket_count,
                     uint32_t *perm, uint32_t *offsets, uint32_t n_payloads, const void *const *pay_in,
                     void *const *pay_out) {
    if (n_payloads > kMkpermMaxPayloads)
        raise(DRJIT_B200_EINVAL, "drjit_b200_call_reduce(): at most %u payload arrays (got %u)!", kMkpermMaxPayloads, n_payloads);
    MkpermExtras ex;
    ex.sort_table = bucket_count <= kSortTableMaxBuckets;
    ex.n_pay = n_payloads;
    for (uint32_t k = 0; k < n_payloads; ++k) {
        ex.pay_in[k] = (const uint32_t *) pay_in[k]; ex.pay_out[k] = (uint32_t *) pay_out[k];
        if (!pay_in[k] || !pay_out[k] || pay_in[k] == pay_out[k])
            raise(DRJIT_B200_EINVAL, "drjit_b200_call_reduce(): payload arrays must be distinct non-null pointers!");
    }
    const uint32_t unique = mkperm_impl(stream, ids, size, std::max(size, 1u), bucket_count, 0, perm, offsets, nullptr, nullptr, ex);
    if (offsets && size && !ex.sort_table) {
        // more buckets than the device-side sort holds: order the rows on the host like the reference
        struct Row { uint32_t id, start, size, unused; };
        Row *rows = reinterpret_cast<Row *>(offsets);
        std::stable_sort(rows, rows + unique, [](const Row &a, const Row &b) { return a.size > b.size; });
    }
    return unique;
}

void mkperm_sharded(cudaStream_t stream, const uint32_t *values, uint32_t size, uint32_t bucket_count,
                    uint32_t index_base, uint32_t *perm, uint32_t *hist_dev) {
    if (size == 0) {
        if (hist_dev && bucket_count)
            DJB_CUDA_CHECK(cudaMemsetAsync(hist_dev, 0, (size_t) bucket_count * 4, stream));
        return;
    }
    mkperm_impl(stream, values, size, size, bucket_count, index_base, perm, nullptr, hist_dev);
}

/// Shard-local permutation of one sorting group that is sharded over the ranks of a communicator,
/// fused with the exchange of the shard histograms (inside the bucket-scan kernel, no library
/// collective, no host round trip): perm = this shard's permutation with entries index_base + local
/// index; hist_dev[b] = this shard's count of bucket b; rank_base_dev[b] = where this rank's keys of
/// bucket b start in the global rank-major (= stable) order; offsets (pinned host, may be NULL) = the
/// table of non-empty buckets of the GLOBAL array, identical on every rank. Returns the unique count
/// when `offsets` is given (after waiting for the table, like the single-GPU call).
uint32_t comm_mkperm(cudaStream_t stream, const Comm *comm, const uint32_t *values, uint32_t size,
                     uint32_t bucket_count, uint32_t index_base, uint32_t *perm, uint32_t *hist_dev,
                     uint32_t *rank_base_dev, uint32_t *offsets) {
    MkpermPeer peer{ comm_ctx(comm), rank_base_dev };
    if ((uint64_t) bucket_count * 4 > kSlotBytes)
        raise(DRJIT_B200_EINVAL, "drjit_b200_comm_mkperm(): at most %u buckets fit the exchange window (got %u)!",
              kSlotBytes / 4, bucket_count);
    return mkperm_impl(stream, values, size, std::max(size, 1u), bucket_count, index_base, perm, offsets, hist_dev, &peer);
}

} // namespace djb
