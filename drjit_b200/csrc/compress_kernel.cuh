/*
 * compress_kernel.cuh -- device side of the mask compaction (host side: compress.cu; geometry
 * sweep: scripts/sweep_compress.cu).
 *
 * Single pass over the mask, static round-robin tile schedule on a cooperative grid, TMA ring
 * for the input, windowed carry between tiles (same scheme as scan_kernel.cuh, "WINDOW").
 * What keeps the kernel off the issue limit (the first version spent 1.05 warp instructions per
 * mask byte, two thirds of them in a per-bit `while (m)` loop, skewed staging addresses and the
 * copy-out: profiles/r1_ncu_full.md, profiles/r1_compress_lines.txt):
 *   * a tile is decoded exactly once, one iteration ahead: the 16-bit masks of the next tile
 *     live in ROWS/2 registers per thread, its warp counts in a 3-deep shared-memory ring. The
 *     TMA stage is released right after that decode, so STAGES tiles are always in flight.
 *   * the tile's count is published as soon as it is decoded; a tile's output offset is the
 *     running sum of the counts of all earlier tiles, which every CTA advances by the G counts
 *     between its previous and its current tile (one or two 8-byte L2 loads per thread, issued
 *     before the decode of the next tile and consumed after it). No look-back, no chain between
 *     CTAs, one __syncthreads per tile.
 *   * per-row ranks come from one shuffle scan per PAIR of rows (two 16-bit counters per word).
 *   * a thread expands its 16-bit mask without a data-dependent loop: two look-ups in a 256-entry
 *     table (byte -> the positions of its set bits as packed nibbles) give a 64-bit nibble
 *     stream; sixteen predicated 16-bit stores with immediate offsets put it into the warp's
 *     staging row (no per-entry address arithmetic), from where the indices leave with fully
 *     coalesced stores; only __syncwarp() is needed.
 */
#pragma once

#include "common.cuh"
#include "tma.cuh"
#include "comm.cuh"

namespace djb {

constexpr uint32_t kCompThreads = 256;
constexpr uint32_t kCompWarps = kCompThreads / 32;
constexpr uint32_t kCompUnit = 16;                                   // mask bytes per load
constexpr uint32_t kCompFetchTid = kCompThreads - 32;                // issues the TMA copies
constexpr uint32_t kCompRowSlots = 32 * kCompUnit;                   // outputs of one warp row

enum : uint32_t { kCInvalid = 0, kCAggregate = 1 };

struct CompressParams {
    const uint8_t *in;
    uint32_t *out;
    uint64_t *state;      // tile descriptors {count << 32 | kCAggregate}, zero on entry
    uint32_t *count_out;  // device-accessible
    uint32_t size, tiles, index_base;
    uint32_t debug;       // read only in -DDRJIT_B200_EXPERIMENTS builds (scripts/sweep_compress.cu): 1 = carry chain disabled
    // PEER instantiations (sharded masks, comm_compress): the thread that learns the shard's count
    // exchanges it with the other ranks through the communicator's scalar cells and writes all W
    // counts, followed by the call's sequence number, into pinned host memory. No second launch, and
    // the host is told when the last tile STARTS its compaction -- it does not wait for the kernel
    // to drain and for the stream's completion signal.
    PeerCtx peer;
    uint32_t *host_counts;  // device view of kMaxPeers + 1 pinned words: counts[0..world), [kMaxPeers] = seq
    uint32_t seq;
};

/// One thread: all-gather of the per-rank counts straight into pinned host memory, then the
/// sequence word the host spins on (also used by the one-thread kernel of an empty shard)
__device__ __forceinline__ void peer_publish_counts(const PeerCtx &c, uint32_t mine, uint32_t *host_counts, uint32_t seq) {
    const uint32_t epoch = peer_begin(c);
    peer_put_u64(c, epoch, mine);
    for (uint32_t src = 0; src < c.world; ++src)
        host_counts[src] = (uint32_t) peer_get_u64(c, epoch, src);
    peer_end(c, epoch);
    __threadfence_system();                         // counts before the sequence word
    *reinterpret_cast<volatile uint32_t *>(host_counts + kMaxPeers) = seq;
}

constexpr uint32_t kCompWindowLoads = 3;                     // carry window: grids of up to 768 CTAs
constexpr uint32_t kComp2RowStride = kCompRowSlots;          // u16 entries per warp staging row
constexpr uint32_t kCompBulkSlots = kCompRowSlots + 8;       // u32 entries per warp staging row, bulk copy-out
                                                             // (shifted by the destination's offset inside 16 bytes)

/// How the indices of a warp row leave the SM
enum : uint32_t {
    kCopyLsu = 0,    // 16-bit staging entries, LDS.U16 + STG.32 per index
    kCopyBulk = 1,   // 32-bit staging entries, one shared -> global bulk copy (TMA engine) per row
    kCopyVec = 2,    // 32-bit staging entries aligned with the destination, LDS.128 + STG.128 per four indices
    kCopyLsuPairs = 3 // kCopyLsu with the staging entries written two at a time (STS.32 of two 16-bit entries)
};

/// Dynamic shared memory of an instantiation: the input stage(s) + the copy-out staging rows
template <uint32_t ROWS, uint32_t STAGES, uint32_t COPY>
constexpr uint32_t compress_smem_bytes() {
    return STAGES * kCompThreads * ROWS * kCompUnit +
           (COPY == kCopyBulk ? kCompWarps * 2 * kCompBulkSlots * 4 :
            COPY == kCopyVec ? kCompWarps * kCompBulkSlots * 4 : kCompWarps * (kComp2RowStride + 8) * 2);
}

/// bit k <- (byte k of the 16-byte unit is non-zero)
__device__ __forceinline__ uint32_t unit_mask16(const uint4 &v) {
    auto top_nibble = [](uint32_t w) -> uint32_t {
        const uint32_t nz = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u; // bit 7 of each non-zero byte
        return nz * 0x00204081u;        // bits 28..31 <- bytes 0..3 (no carries: all partial products are disjoint)
    };
    return (top_nibble(v.x) >> 28) | ((top_nibble(v.y) >> 24) & 0xf0u) |
           ((top_nibble(v.z) >> 20) & 0xf00u) | ((top_nibble(v.w) >> 16) & 0xf000u);
}

__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"((uint16_t) v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
/// (a & b) | c in one LOP3
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

/// COPY: kCopyLsu / kCopyBulk. BASE512: p.index_base is a multiple of 512, so that the index of an
/// entry is the OR of (row base | lane * 16 | position nibble) -- no add per entry (kCopyBulk only).
template <uint32_t ROWS, uint32_t STAGES, uint32_t MIN_CTAS, uint32_t COPY = kCopyLsu, bool BASE512 = true, bool PEER = false>
__global__ void __launch_bounds__(kCompThreads, MIN_CTAS)
compress_kernel(const CompressParams p) {
    static_assert(ROWS % 2 == 0, "rows are ranked in pairs");
    constexpr uint32_t TILE = kCompThreads * ROWS * kCompUnit;
    constexpr uint32_t PAIRS = ROWS / 2;
    constexpr bool STAGED = STAGES > 0;
    constexpr uint32_t NS = STAGED ? STAGES : 1;
    static_assert(TILE <= 65536, "tile-local offsets are staged as 16-bit values");
    extern __shared__ __align__(128) uint8_t stage_mem[];       // STAGES x TILE, then the copy-out staging rows
    uint8_t *const row_stage_mem = stage_mem + (size_t) STAGES * TILE;
    __shared__ uint32_t lut[256];           // byte -> positions of its set bits, one nibble each, ascending
    __shared__ uint64_t full_bar[NS];
    __shared__ uint32_t wcnt[4][kCompWarps];    // per-warp counts of tiles it-1 .. it+2
    __shared__ uint32_t win_cnt[2][kCompWarps]; // per-warp partial sums of the carry window

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t size = p.size;
    const bool aligned = (((uintptr_t) p.in) & 15u) == 0;

    {   // expansion table
        uint32_t e = 0, n = 0;
        #pragma unroll
        for (uint32_t b = 0; b < 8; ++b)
            if ((tid >> b) & 1u) { e |= b << (4 * n); ++n; }
        lut[tid] = e;
    }

    uint64_t policy = 0;
    auto tile_of = [&](uint32_t k) -> uint64_t { return (uint64_t) blockIdx.x + (uint64_t) k * gridDim.x; };
    auto tile_is_staged = [&](uint64_t tile) -> bool {
        return STAGED && tile < p.tiles && (tile + 1) * TILE <= size;
    };
    auto issue = [&](uint32_t s, uint64_t tile) {
        if (tile_is_staged(tile)) {
            mbar_expect_tx(&full_bar[s], TILE);
            bulk_load(stage_mem + (size_t) s * TILE, p.in + tile * TILE, TILE, &full_bar[s], policy);
        }
    };
    if constexpr (STAGED) {
        if (tid == kCompFetchTid) {
            #pragma unroll
            for (uint32_t s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
            fence_proxy_async();
            policy = policy_evict_first();
            #pragma unroll
            for (uint32_t s = 0; s < STAGES; ++s) issue(s, tile_of(s));
        }
    }
    __syncthreads();

    // Masks of this CTA's k-th tile -> mk (two 16-bit masks per word); the warp's count goes
    // to the ring. Rows past the end of the array decode as empty.
    auto decode = [&](uint32_t k, uint32_t (&mk)[PAIRS]) {
        const uint64_t tile64 = tile_of(k);
        #pragma unroll
        for (uint32_t i = 0; i < PAIRS; ++i) mk[i] = 0;
        if (tile64 >= p.tiles)
            return;
        if (tile_is_staged(tile64)) {
            const uint32_t s = k % NS;
            mbar_wait(&full_bar[s], (k / NS) & 1u);
            const uint8_t *src = stage_mem + (size_t) s * TILE;
            #pragma unroll
            for (uint32_t r = 0; r < ROWS; ++r) {
                const uint4 v = lds128(src + ((warp * ROWS + r) * 32 + lane) * kCompUnit);
                mk[r / 2] |= unit_mask16(v) << (16 * (r & 1u));
            }
        } else {
            const uint64_t tile_base = tile64 * TILE;
            #pragma unroll
            for (uint32_t r = 0; r < ROWS; ++r) {
                const uint64_t s0 = tile_base + (uint64_t) (((warp * ROWS + r) * 32 + lane) * kCompUnit);
                uint32_t m = 0;
                if (s0 < size) {
                    if (aligned && s0 + kCompUnit <= size) {
                        const Vec16<uint32_t> v = ld_stream<uint32_t>(p.in + s0);
                        m = unit_mask16(*reinterpret_cast<const uint4 *>(&v));
                    } else {
                        #pragma unroll
                        for (uint32_t e = 0; e < kCompUnit; ++e)
                            if (s0 + e < size && p.in[s0 + e] != 0)
                                m |= 1u << e;
                    }
                }
                mk[r / 2] |= m << (16 * (r & 1u));
            }
        }
        uint32_t c = 0;
        #pragma unroll
        for (uint32_t i = 0; i < PAIRS; ++i) c += __popc(mk[i]);
        c = __reduce_add_sync(kFullMask, c);
        if (lane == 0) wcnt[k % 4][warp] = c;
    };
    // warp 0, after the barrier that follows decode(k): the tile's count becomes visible to
    // the look-backs of other CTAs one whole iteration before the tile itself is compacted
    auto publish_aggregate = [&](uint32_t k) {
        const uint64_t t = tile_of(k);
        if (t >= p.tiles)
            return;
        uint32_t c = lane < kCompWarps ? wcnt[k % 4][lane] : 0u;
        c = __reduce_add_sync(kFullMask, c);
        if (lane == 0)
            st_relaxed_u64(p.state + t, ((uint64_t) c << 32) | kCAggregate);
    };

    // Tiles are decoded (and their counts published) two iterations before they are compacted, the
    // same slack the scan's three-stage ring gives its aggregates: the carry windows below then
    // find every count in place even when CTAs drift apart by an iteration.
    uint32_t cur[PAIRS], nxt[PAIRS], nxt2[PAIRS];
    uint32_t carry = 0;                 // selected entries in all tiles before the current one
    decode(0, cur);
    __syncthreads();                                    // stage 0 is free, counts visible
    if constexpr (STAGED) {
        if (tid == kCompFetchTid) issue(0, tile_of(NS));
    }
    decode(1, nxt);
    __syncthreads();
    if constexpr (STAGED) {
        if (tid == kCompFetchTid) issue(1 % NS, tile_of(1 + NS));
    }
    if (warp == 0) { publish_aggregate(0); publish_aggregate(1); }

    const uint32_t lane16 = lane * kCompUnit;
    const uint32_t stg_addr = smem_addr(row_stage_mem) +
                              warp * (COPY == kCopyBulk ? 2 * kCompBulkSlots * 4 :
                                      COPY == kCopyVec ? kCompBulkSlots * 4 : (kComp2RowStride + 8) * 2);
    uint32_t bulk_par = 0;              // kCopyBulk: which of the warp's two staging rows is written next

    for (uint32_t it = 0;; ++it) {
        const uint64_t tile64 = tile_of(it);
        if (tile64 >= p.tiles)
            break;
        const uint32_t tile = (uint32_t) tile64;

        // ---- carry window: counts of the tiles between this CTA's previous tile and this one ----
        // (scan_kernel.cuh, "WINDOW")
        uint64_t wd[kCompWindowLoads];
        const uint32_t win_lo = it == 0 ? 0u : tile - gridDim.x, win_n = (DJB_DEBUG(p.debug) & 1u) ? 0u : tile - win_lo;
        #pragma unroll
        for (uint32_t j = 0; j < kCompWindowLoads; ++j) {
            const uint32_t o = j * kCompThreads + tid;
            wd[j] = o < win_n ? ld_relaxed_u64(p.state + win_lo + o) : (uint64_t) kCAggregate;
        }
        // ---- decode the tile after the next one ----------------------------------------------
        decode(it + 2, nxt2);
        {
            uint32_t w = 0;
            #pragma unroll
            for (uint32_t j = 0; j < kCompWindowLoads; ++j) {
                // a CTA that runs behind: poll, backing off. On sparse masks an iteration is as short as a
                // memory round trip and the grid advances in step with its slowest tile load: a third of
                // all issued instructions were 20 ns polls (1 % density, profiles/r5g_ncu_compress01.md).
                // They did not cost time (A/B: 0.381 against 0.379 ms, profiles/r5h_compress_poll_ab.txt)
                // -- the issue slots were free -- but there is no reason to burn them.
                for (uint32_t ns = 32; (uint32_t) wd[j] == kCInvalid; ns = ns < 256 ? ns * 2 : ns) {
                    __nanosleep(ns);
                    wd[j] = ld_relaxed_u64(p.state + win_lo + j * kCompThreads + tid);
                }
                w += (uint32_t) (wd[j] >> 32);
            }
            w = __reduce_add_sync(kFullMask, w);
            if (lane == 0) win_cnt[it & 1u][warp] = w;
        }
        __syncthreads();            // window sums, counts of tile it+2, stage of tile it+2 free
        if constexpr (STAGED) {
            if (tid == kCompFetchTid) issue((it + 2) % NS, tile_of(it + 2 + STAGES));
        }
        if (warp == 0) publish_aggregate(it + 2);
        #pragma unroll
        for (uint32_t w = 0; w < kCompWarps; ++w)
            carry += win_cnt[it & 1u][w];
        if (DJB_DEBUG(p.debug) & 1u) carry = tile * (TILE / 2);       // (keeps the write stream spread out)
        if (tile == p.tiles - 1 && tid == 0) {
            uint32_t ttotal = 0;
            #pragma unroll
            for (uint32_t w = 0; w < kCompWarps; ++w) ttotal += wcnt[it % 4][w];
            *p.count_out = carry + ttotal;
            if constexpr (PEER)
                peer_publish_counts(p.peer, carry + ttotal, p.host_counts, p.seq);
        }

        // ---- compaction of tile `it` -----------------------------------------------------------
        uint32_t wprefix = 0;
        #pragma unroll
        for (uint32_t w = 0; w < kCompWarps; ++w)
            if (w < warp) wprefix += wcnt[it % 4][w];
        uint32_t *dst = p.out + carry + wprefix;
        const uint32_t idx_warp = p.index_base + (uint32_t) (tile64 * TILE) + warp * ROWS * kCompRowSlots;

        #pragma unroll
        for (uint32_t i = 0; i < PAIRS; ++i) {
            // ranks of two rows at once: 16-bit counters (a row selects at most 512 entries)
            const uint32_t c2 = __popc(cur[i] & 0xffffu) | (__popc(cur[i] >> 16) << 16);
            uint32_t v = c2;
            #pragma unroll
            for (uint32_t d = 1; d < 32; d <<= 1) {
                const uint32_t t = shfl_up(v, d);
                if (lane >= d) v += t;
            }
            const uint32_t tot2 = shfl_idx(v, 31), ex2 = v - c2;
            #pragma unroll
            for (uint32_t h = 0; h < 2; ++h) {
                const uint32_t m = (cur[i] >> (16 * h)) & 0xffffu;
                const uint32_t c = (c2 >> (16 * h)) & 0xffffu, r = (ex2 >> (16 * h)) & 0xffffu,
                               n = (tot2 >> (16 * h)) & 0xffffu;
                // 64-bit nibble stream: positions of the set bits of m, ascending
                const uint32_t lo = lut[m & 0xffu], hi = lut[m >> 8] + 0x88888888u;
                const uint64_t q = (uint64_t) lo | ((uint64_t) hi << (4 * __popc(m & 0xffu)));
                const uint32_t idx_row = idx_warp + (2 * i + h) * kCompRowSlots;
                if constexpr (COPY == kCopyBulk || COPY == kCopyVec) {
                    // 32-bit staging entries hold the final index. Staging slot s of the row corresponds
                    // to dst[s - a], a = position of dst inside its 16-byte line, so that aligned groups of
                    // four slots are aligned groups of four output words:
                    //   kCopyVec : LDS.128 + STG.128 per four indices (3.5 LSU instructions per 128 indices
                    //              instead of 8, no per-index add), scalar stores for up to 3 + 3 entries at
                    //              the ragged ends;
                    //   kCopyBulk: the aligned middle leaves with ONE shared -> global bulk copy issued by
                    //              lane 0; two staging rows per warp so that the copy of row k reads its row
                    //              while row k + 1 is being staged (measured slower: profiles/r4b_*).
                    const uint32_t a = ((uint32_t) (uintptr_t) dst >> 2) & 3u;
                    uint32_t buf = stg_addr;
                    if constexpr (COPY == kCopyBulk) {
                        buf += bulk_par * (kCompBulkSlots * 4);
                        bulk_par ^= 1u;
                        if (lane == 0) bulk_wait_read<1>();     // the copy issued two rows ago has read this row
                        __syncwarp();
                    }
                    const uint32_t base = BASE512 ? (idx_row | lane16) : idx_row + lane16;
                    uint32_t wa = buf + 4 * (a + r);
                    asm volatile("" : "+r"(wa));
                    auto entry = [&](uint32_t nib) -> uint32_t {
                        return BASE512 ? and_or(nib, 15u, base) : (nib & 15u) + base;
                    };
                    if (n > 64) {
                        #pragma unroll
                        for (uint32_t j = 0; j < kCompUnit; ++j)
                            if (j < c)
                                sts_u32(wa + 4 * j, entry((uint32_t) (q >> (4 * j))));
                    } else {
                        const uint32_t cmax = __reduce_max_sync(kFullMask, c);
                        uint64_t qq = q;
                        for (uint32_t j = 0; j < cmax; ++j, qq >>= 4)
                            if (j < c)
                                sts_u32(wa + 4 * j, entry((uint32_t) qq));
                    }
                    if constexpr (COPY == kCopyBulk)
                        fence_proxy_async();                    // staged entries -> visible to the bulk copy engine
                    __syncwarp();
                    const uint32_t end = a + n, s0 = (a + 3u) & ~3u, s1 = end & ~3u;
                    const bool body = s1 > s0;
                    if (lane < 8 && lane >= a && lane < end && (!body || lane < s0))
                        dst[lane - a] = lds_u32(buf + 4 * lane);
                    if (body) {
                        const uint32_t t = s1 + lane;
                        if (lane < 3 && t < end)
                            dst[t - a] = lds_u32(buf + 4 * t);
                        if constexpr (COPY == kCopyBulk) {
                            if (lane == 0)
                                bulk_store(dst - a + s0, buf + 4 * s0, (s1 - s0) * 4);
                        } else {
                            // groups of four slots [s0 / 4, s1 / 4): lane takes group s0 / 4 + lane + 32 t
                            const uint32_t g1 = s1 >> 2;
                            uint32_t g = (s0 >> 2) + lane;
                            uint4 *o = reinterpret_cast<uint4 *>(dst - a) + g;      // (dst - a is 16-byte aligned)
                            uint32_t sa = buf + 16 * g;
                            for (; g < g1; g += 32, o += 32, sa += 512) {
                                uint4 v;
                                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa) : "memory");
                                asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                                             :: "l"(o), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                            }
                        }
                    }
                    if constexpr (COPY == kCopyBulk) {
                        if (lane == 0) bulk_commit();
                    } else {
                        __syncwarp();                           // row copied out before the next one is staged
                    }
                    dst += n;
                    continue;
                }
                uint32_t wa = stg_addr + 2 * r;
                asm volatile("" : "+r"(wa));      // (keeps ptxas from re-deriving the address per store)
                if (COPY == kCopyLsuPairs && n > 64) {
                    // Two entries per store: half as many shared-memory store instructions (and bank
                    // conflict wavefronts) as the entry-wise loop below. A lane whose start r is odd
                    // keeps its first entry for a 16-bit store and pairs up the rest from r + 1; a
                    // lane with an odd number of paired entries writes one garbage half into the slot
                    // after its last entry -- that slot is the (odd) start of the next non-empty lane,
                    // whose 16-bit store comes AFTER all pair stores and puts the right entry there.
                    const uint32_t odd = r & 1u;
                    const uint32_t first = and_or((uint32_t) q, 15u, lane16);
                    const uint32_t qlo = odd ? (uint32_t) (q >> 4) : (uint32_t) q,
                                   qhi = odd ? (uint32_t) (q >> 36) : (uint32_t) (q >> 32);
                    const uint32_t cp = c - odd * (c != 0);                 // entries that go out in pairs
                    uint32_t wp = wa + 2 * odd;                             // 4-byte aligned
                    asm volatile("" : "+r"(wp));
                    const uint32_t lane2 = lane16 * 0x10001u;
                    #pragma unroll
                    for (uint32_t j = 0; j < kCompUnit / 2; ++j) {
                        // byte j of the nibble stream = entries 2j, 2j+1 -> spread to two 16-bit fields
                        const uint32_t xb = __byte_perm(j < 4 ? qlo : qhi, 0u, 0x4440u | (j & 3u));   // (selector 4 = a zero byte)
                        if (2 * j < cp)
                            sts_u32(wp + 4 * j, and_or(xb * 0x1001u, 0x000f000fu, lane2));
                    }
                    __syncwarp();
                    if (odd && c != 0)
                        sts_u16(wa, first);
                } else if (n > 64) {
                    #pragma unroll
                    for (uint32_t j = 0; j < kCompUnit; ++j)
                        if (j < c)
                            sts_u16(wa + 2 * j, and_or((uint32_t) (q >> (4 * j)), 15u, lane16));
                } else {
                    // sparse row (n is warp-uniform): only as many rounds as the fullest lane needs
                    const uint32_t cmax = __reduce_max_sync(kFullMask, c);
                    uint64_t qq = q;
                    for (uint32_t j = 0; j < cmax; ++j, qq >>= 4)
                        if (j < c)
                            sts_u16(wa + 2 * j, and_or((uint32_t) qq, 15u, lane16));
                }
                __syncwarp();
                // coalesced copy-out, 128 slots per round (a two-slots-per-lane LDS.32/STG.64
                // variant was 18 % slower: profiles/r1f_sweep_compress_paired_copyout.txt)
                {   // (pointers and the remaining count are stepped explicitly: immediates only inside)
                    uint32_t *o = dst + lane;
                    uint32_t sa = stg_addr + 2 * lane;
                    int32_t rem = (int32_t) n - (int32_t) lane;             // slots left for this lane: rem > 32 t
                    int32_t left = (int32_t) n;                              // slots left for the warp (uniform)
                    for (; left > 96; left -= 128, rem -= 128, o += 128, sa += 256) {
                        #pragma unroll
                        for (uint32_t t = 0; t < 4; ++t)
                            if (rem > (int32_t) (t * 32))
                                o[t * 32] = idx_row + lds_u16(sa + t * 64);
                    }
                    for (; left > 0; left -= 32, rem -= 32, o += 32, sa += 64)   // short tail: 32 slots per round
                        if (rem > 0)
                            *o = idx_row + lds_u16(sa);
                }
                __syncwarp();
                dst += n;
            }
        }
        #pragma unroll
        for (uint32_t i = 0; i < PAIRS; ++i) { cur[i] = nxt[i]; nxt[i] = nxt2[i]; }
    }
    if constexpr (COPY == kCopyBulk) {
        if (lane == 0) bulk_wait_read<0>();     // the staging rows must outlive the copies that read them
    }
}

} // namespace djb
