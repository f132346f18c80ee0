/*
 * compress_kernel.cuh -- device side of the mask compaction (see compress.cu for the design).
 * Host-side dispatch: compress.cu; geometry sweep: scripts/sweep_compress.cu.
 */
#pragma once

#include "common.cuh"
#include "tma.cuh"

namespace djb {

constexpr uint32_t kCompThreads = 256;
constexpr uint32_t kCompWarps = kCompThreads / 32;
constexpr uint32_t kCompUnit = 16;                                   // mask bytes per load
constexpr uint32_t kCompFetchTid = kCompThreads - 32;                // issues the TMA copies
constexpr uint32_t kCompLookbackLoads = 8;                           // look-back window = 256 tiles
constexpr uint32_t kCompRowSlots = 32 * kCompUnit;                   // outputs of one warp row
constexpr uint32_t kCompRowStride = kCompRowSlots + (kCompRowSlots >> 6) * 2 + 8;  // skewed, u16 entries

enum : uint32_t { kCInvalid = 0, kCAggregate = 1, kCPrefix = 2 };

struct CompressParams {
    const uint8_t *in;
    uint32_t *out;
    uint64_t *state;      // tile descriptors {count << 32 | status}, zero on entry
    uint32_t *count_out;  // device-accessible
    uint32_t size, tiles, index_base;
};

/// One bit per non-zero byte of a 32-bit word (bit k <- byte k)
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t w) {
    const uint32_t nz = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u; // bit 7 of each non-zero byte
    return (((nz >> 7) * 0x01020408u) >> 24) & 0xfu;
}

/// Staging slot (16-bit entries) -> halfword index in shared memory. One padding word per
/// 32 words keeps the runs written by different lanes on different banks even when every
/// lane writes a full 16-entry run ("compress").
__device__ __forceinline__ uint32_t skew(uint32_t slot) { return slot + ((slot >> 6) << 1); }

/// ROWS 16-byte units per thread: tile = 256 * ROWS * 16 mask bytes (ROWS = 8: 32768).
/// STAGES > 0: 16-byte aligned masks, full tiles arrive through the TMA ring; STAGES == 0:
/// direct loads (unaligned masks).
template <uint32_t ROWS, uint32_t STAGES, uint32_t MIN_CTAS>
__global__ void __launch_bounds__(kCompThreads, MIN_CTAS)
compress_kernel(const CompressParams p) {
    constexpr uint32_t TILE = kCompThreads * ROWS * kCompUnit;
    constexpr bool STAGED = STAGES > 0;
    constexpr bool EARLY = STAGES >= 2;     // early tile counts, see scan_kernel.cuh
    static_assert(TILE <= 65536, "tile-local offsets are staged as 16-bit values");
    extern __shared__ __align__(128) uint8_t stage_mem[];       // STAGES x TILE
    __shared__ uint16_t row_stage[kCompWarps][kCompRowStride];
    __shared__ uint64_t full_bar[STAGED ? STAGES : 1];
    __shared__ uint32_t warp_cnt[kCompWarps];
    __shared__ uint32_t early_cnt[kCompWarps];
    __shared__ uint32_t base_smem;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t size = p.size;

    // static round-robin tile schedule + cooperative launch, see scan_kernel.cuh
    uint64_t policy = 0;
    auto tile_is_staged = [&](uint32_t tile) -> bool {
        return tile < p.tiles && (uint64_t) (tile + 1) * TILE <= size;
    };
    auto issue = [&](uint32_t s, uint32_t tile) {
        if (tile_is_staged(tile)) {
            mbar_expect_tx(&full_bar[s], TILE);
            bulk_load(stage_mem + (size_t) s * TILE, p.in + (uint64_t) tile * TILE, TILE, &full_bar[s], policy);
        }
    };
    if constexpr (STAGED) {
        if (tid == kCompFetchTid) {
            #pragma unroll
            for (uint32_t s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
            fence_proxy_async();
            policy = policy_evict_first();
            #pragma unroll
            for (uint32_t s = 0; s < STAGES; ++s) issue(s, blockIdx.x + s * gridDim.x);
        }
        __syncthreads();
    }

    // EARLY: count the k-th tile of this CTA straight from its stage as soon as it has landed
    // and publish the count STAGES-1 iterations before the tile itself is compacted
    auto early_count = [&](uint32_t k) -> bool {
        const uint64_t t64 = (uint64_t) blockIdx.x + (uint64_t) k * gridDim.x;
        if (t64 >= p.tiles || !tile_is_staged((uint32_t) t64))
            return false;
        constexpr uint32_t NS = STAGED ? STAGES : 1;
        const uint32_t s = k % NS;
        mbar_wait(&full_bar[s], (k / NS) & 1u);
        const uint8_t *src = stage_mem + (size_t) s * TILE;
        uint32_t c = 0;
        #pragma unroll
        for (uint32_t r = 0; r < ROWS; ++r) {
            const uint4 v = lds128(src + ((warp * ROWS + r) * 32 + lane) * kCompUnit);
            c += __popc(nonzero_nibble(v.x) | (nonzero_nibble(v.y) << 4) |
                        (nonzero_nibble(v.z) << 8) | (nonzero_nibble(v.w) << 12));
        }
        c = __reduce_add_sync(kFullMask, c);
        if (lane == 0) early_cnt[warp] = c;
        return true;
    };
    auto early_publish = [&](uint32_t k) {
        uint32_t c = lane < kCompWarps ? early_cnt[lane] : 0u;
        c = __reduce_add_sync(kFullMask, c);
        const uint32_t t = blockIdx.x + k * gridDim.x;
        if (lane == 0 && t != 0)
            st_relaxed_u64(p.state + t, ((uint64_t) c << 32) | kCAggregate);
    };
    if constexpr (EARLY) {
        for (uint32_t k = 0; k + 1 < STAGES; ++k) {
            const bool did = early_count(k);
            __syncthreads();
            if (warp == 0 && did) early_publish(k);
            __syncthreads();
        }
    }

    for (uint32_t it = 0;; ++it) {
        const uint64_t tile64 = (uint64_t) blockIdx.x + (uint64_t) it * gridDim.x;
        if (tile64 >= p.tiles)
            break;
        const uint32_t tile = (uint32_t) tile64;
        const uint32_t stage = STAGED ? it % STAGES : 0;
        const uint64_t tile_base = (uint64_t) tile * TILE;
        const bool staged = STAGED && tile_is_staged(tile);

        // ---- load, byte flags -> bit masks ------------------------------------------
        uint32_t mask[ROWS];
        bool early_done = false;
        if constexpr (EARLY)
            early_done = early_count(it + STAGES - 1);
        if constexpr (STAGED) {
            if (staged) {
                mbar_wait(&full_bar[stage], (it / STAGES) & 1u);
                const uint8_t *src = stage_mem + (size_t) stage * TILE;
                #pragma unroll
                for (uint32_t k = 0; k < ROWS; ++k) {
                    const uint4 v = lds128(src + ((warp * ROWS + k) * 32 + lane) * kCompUnit);
                    mask[k] = nonzero_nibble(v.x) | (nonzero_nibble(v.y) << 4) |
                              (nonzero_nibble(v.z) << 8) | (nonzero_nibble(v.w) << 12);
                }
            }
            __syncthreads();                         // stage is free again
            if (tid == kCompFetchTid) {
                const uint64_t next = tile64 + (uint64_t) STAGES * gridDim.x;
                if (next < p.tiles) issue(stage, (uint32_t) next);
            }
            if constexpr (EARLY) {
                if (warp == 0 && early_done) early_publish(it + STAGES - 1);
            }
        }
        if (!staged) {
            const bool aligned = (((uintptr_t) p.in) & 15u) == 0;
            #pragma unroll
            for (uint32_t k = 0; k < ROWS; ++k) {
                const uint64_t s0 = tile_base + (uint64_t) (((warp * ROWS + k) * 32 + lane) * kCompUnit);
                uint32_t m = 0;
                if (s0 < size) {
                    if (aligned && s0 + kCompUnit <= size) {
                        const Vec16<uint32_t> v = ld_stream<uint32_t>(p.in + s0);
                        #pragma unroll
                        for (uint32_t j = 0; j < 4; ++j)
                            m |= nonzero_nibble(v.v[j]) << (4 * j);
                    } else {
                        #pragma unroll
                        for (uint32_t e = 0; e < kCompUnit; ++e)
                            if (s0 + e < size && p.in[s0 + e] != 0)
                                m |= 1u << e;
                    }
                }
                mask[k] = m;
            }
        }

        // ---- ranks inside the warp (warp-contiguous item order: row-major, then lane) ------
        uint32_t rank[ROWS], row_total[ROWS], wtotal = 0;
        #pragma unroll
        for (uint32_t k = 0; k < ROWS; ++k) {
            const uint32_t c = __popc(mask[k]);
            uint32_t v = c;
            #pragma unroll
            for (uint32_t d = 1; d < 32; d <<= 1) {
                const uint32_t t = shfl_up(v, d);
                if (lane >= d) v += t;
            }
            rank[k] = v - c;                      // exclusive, inside the row
            row_total[k] = shfl_idx(v, 31);
            wtotal += row_total[k];
        }
        if (lane == 0)
            warp_cnt[warp] = wtotal;
        __syncthreads();

        uint32_t wprefix = 0, ttotal = 0;
        #pragma unroll
        for (uint32_t w = 0; w < kCompWarps; ++w) {
            if (w == warp) wprefix = ttotal;
            ttotal += warp_cnt[w];
        }

        // ---- look-back for the tile's first output slot (warp 0) -----------------------
        if (warp == 0) {
            uint32_t excl = 0;
            if (tile == 0) {
                if (lane == 0)
                    st_relaxed_u64(p.state, ((uint64_t) ttotal << 32) | kCPrefix);
            } else {
                if (lane == 0 && !(EARLY && staged))
                    st_relaxed_u64(p.state + tile, ((uint64_t) ttotal << 32) | kCAggregate);
                int32_t pred = (int32_t) tile - 1 - (int32_t) lane;
                auto consume = [&](int32_t first, uint64_t w) -> bool {
                    while (__any_sync(kFullMask, (uint32_t) w == kCInvalid)) {
                        __nanosleep(20);
                        if (first >= 0) w = ld_relaxed_u64(p.state + first);
                    }
                    const uint32_t done = __ballot_sync(kFullMask, (uint32_t) w == kCPrefix);
                    const uint32_t stop = done ? (uint32_t) __ffs(done) - 1 : 31u;
                    excl += __reduce_add_sync(kFullMask, lane <= stop ? (uint32_t) (w >> 32) : 0u);
                    return done != 0;
                };
                while (true) {
                    // 8 windows of 32 descriptors per round, all loads in flight together; lanes
                    // past the array start behave like a finished tile with count 0
                    uint64_t w[kCompLookbackLoads];
                    #pragma unroll
                    for (uint32_t j = 0; j < kCompLookbackLoads; ++j) {
                        const int32_t idx = pred - 32 * (int32_t) j;
                        w[j] = idx >= 0 ? ld_relaxed_u64(p.state + idx) : (uint64_t) kCPrefix;
                    }
                    bool found = false;
                    #pragma unroll
                    for (uint32_t j = 0; j < kCompLookbackLoads; ++j) {
                        if (!found && consume(pred - 32 * (int32_t) j, w[j]))
                            found = true;
                    }
                    if (found) break;
                    pred -= 32 * (int32_t) kCompLookbackLoads;
                }
                if (lane == 0)
                    st_relaxed_u64(p.state + tile, ((uint64_t) (excl + ttotal) << 32) | kCPrefix);
            }
            if (lane == 0) {
                base_smem = excl;
                if (tile == p.tiles - 1)
                    *p.count_out = excl + ttotal;
            }
        }
        __syncthreads();

        // ---- per warp: expand each row into the private staging row, stream it out ----------
        uint32_t *dst = p.out + base_smem + wprefix;
        const uint32_t idx0 = p.index_base + (uint32_t) tile_base;
        uint16_t *stg = row_stage[warp];
        #pragma unroll
        for (uint32_t k = 0; k < ROWS; ++k) {
            const uint32_t local0 = ((warp * ROWS + k) * 32 + lane) * kCompUnit;
            uint32_t m = mask[k], r = rank[k];
            while (m) {
                const uint32_t b = (uint32_t) __ffs(m) - 1;
                m &= m - 1;
                stg[skew(r)] = (uint16_t) (local0 + b);
                ++r;
            }
            __syncwarp();
            const uint32_t n = row_total[k];
            for (uint32_t s = lane; s < n; s += 32)
                dst[s] = idx0 + stg[skew(s)];
            __syncwarp();
            dst += n;
        }
    }
}

} // namespace djb
