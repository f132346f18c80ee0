/*
 * compress.cu -- mask compaction: indices of the non-zero bytes of a mask, ascending.
 *
 * Replaces CUDAThreadState::compress (ext/drjit-core/src/cuda_ts.cpp:683-763) and the kernels
 * compress_small / compress_large / compress_large_init (resources/compress.cuh:23-156).
 *
 * Design. Same single-pass skeleton as prefix_reduce.cu (ticketed persistent tiles, decoupled
 * look-back on packed 64-bit descriptors), specialised for 1-byte flags:
 *  - a tile is 32768 mask bytes (8192 for small masks); each thread loads 8 x 16 bytes
 *    (LDG.128, warp-striped, all in flight at once) and turns each 16-byte unit into a 16-bit
 *    mask with three integer ops per word, so ranks come from popc instead of a 17-step scalar
 *    scan per thread (compress.cuh:101-109);
 *  - the tile-local offsets of the selected items are first compacted into a (bank-skewed)
 *    16-bit shared-memory staging buffer and then streamed out as indices with fully coalesced
 *    stores; the reference issues scattered 4-byte stores straight to global memory
 *    (compress.cuh:151-154);
 *  - the mask is never written (the reference zero-pads the caller's buffer, cuda_ts.cpp:746-748);
 *    the ragged tail is bounds-checked instead;
 *  - the count goes to a device-mapped pinned word, one launch + one memset in total.
 */
#include "common.cuh"
#include "runtime.h"

namespace djb {

constexpr uint32_t kCompThreads = 256;
constexpr uint32_t kCompWarps = kCompThreads / 32;
constexpr uint32_t kCompUnit = 16;                                   // mask bytes per load

enum : uint32_t { kCInvalid = 0, kCAggregate = 1, kCPrefix = 2 };

struct CompressParams {
    const uint8_t *in;
    uint32_t *out;
    uint64_t *state;      // tile descriptors {count << 32 | status}
    uint32_t *ticket;
    uint32_t *count_out;  // device-accessible
    uint32_t size, tiles, index_base;
    uint8_t vec;
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
/// Large tiles bound the tile rate that the look-back has to follow (see prefix_reduce.cu).
template <uint32_t ROWS>
__global__ void __launch_bounds__(kCompThreads)
compress_kernel(const CompressParams p) {
    constexpr uint32_t TILE = kCompThreads * ROWS * kCompUnit;
    static_assert(TILE <= 65536, "tile-local offsets are staged as 16-bit values");
    extern __shared__ uint16_t staged[];      // skew(TILE) entries
    __shared__ uint32_t warp_cnt[kCompWarps];
    __shared__ uint32_t tile_smem, base_smem;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t size = p.size;

    while (true) {
        if (tid == 0)
            tile_smem = atomicAdd(p.ticket, 1u);
        __syncthreads();
        const uint32_t tile = tile_smem;
        if (tile >= p.tiles)
            break;
        const uint64_t tile_base = (uint64_t) tile * TILE;

        // ---- load, byte flags -> bit masks ------------------------------------------
        uint32_t mask[ROWS];
        if (p.vec && tile_base + TILE <= size) {
            // full tile: issue all loads first
            Vec16<uint32_t> v[ROWS];
            #pragma unroll
            for (uint32_t k = 0; k < ROWS; ++k)
                v[k] = ld_stream<uint32_t>(p.in + tile_base + ((warp * ROWS + k) * 32 + lane) * kCompUnit);
            #pragma unroll
            for (uint32_t k = 0; k < ROWS; ++k) {
                uint32_t m = 0;
                #pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                    m |= nonzero_nibble(v[k].v[j]) << (4 * j);
                mask[k] = m;
            }
        } else {
            #pragma unroll
            for (uint32_t k = 0; k < ROWS; ++k) {
                const uint64_t s0 = tile_base + (uint64_t) (((warp * ROWS + k) * 32 + lane) * kCompUnit);
                uint32_t m = 0;
                if (s0 < size) {
                    if (p.vec && s0 + kCompUnit <= size) {
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
        uint32_t rank[ROWS], wtotal = 0;
        #pragma unroll
        for (uint32_t k = 0; k < ROWS; ++k) {
            const uint32_t c = __popc(mask[k]);
            uint32_t v = c;
            #pragma unroll
            for (uint32_t d = 1; d < 32; d <<= 1) {
                const uint32_t t = shfl_up(v, d);
                if (lane >= d) v += t;
            }
            rank[k] = wtotal + v - c;
            wtotal += shfl_idx(v, 31);
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
                if (lane == 0)
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
                    // two windows of 32 descriptors per round; lanes past the array start
                    // behave like a finished tile with count 0
                    uint64_t w0 = kCPrefix, w1 = kCPrefix;
                    if (pred >= 0) w0 = ld_relaxed_u64(p.state + pred);
                    if (pred >= 32) w1 = ld_relaxed_u64(p.state + pred - 32);
                    if (consume(pred, w0)) break;
                    if (consume(pred - 32, w1)) break;
                    pred -= 64;
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

        // ---- compact tile-local offsets of the selected items into shared memory ---------
        #pragma unroll
        for (uint32_t k = 0; k < ROWS; ++k) {
            const uint32_t local0 = ((warp * ROWS + k) * 32 + lane) * kCompUnit;
            uint32_t m = mask[k], r = wprefix + rank[k];
            while (m) {
                const uint32_t b = (uint32_t) __ffs(m) - 1;
                m &= m - 1;
                staged[skew(r)] = (uint16_t) (local0 + b);
                ++r;
            }
        }
        __syncthreads();

        // ---- coalesced write-out -------------------------------------------------------------
        uint32_t *dst = p.out + base_smem;
        const uint32_t idx0 = p.index_base + (uint32_t) tile_base;
        for (uint32_t s = tid; s < ttotal; s += kCompThreads)
            dst[s] = idx0 + staged[skew(s)];
    }
}

constexpr uint32_t kCompRowsBig = 8, kCompRowsSmall = 2;

template <uint32_t ROWS>
static void launch_compress(cudaStream_t stream, CompressParams &p, Scratch &scratch) {
    constexpr uint32_t TILE = kCompThreads * ROWS * kCompUnit;
    const DeviceProps &dev = device_props();
    const uint32_t smem = (TILE + (TILE >> 6) * 2 + 64) * sizeof(uint16_t);
    static int occupancy = 0;
    if (occupancy == 0) {
        if (smem > 48 * 1024)
            DJB_CUDA_CHECK(cudaFuncSetAttribute(compress_kernel<ROWS>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        DJB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, compress_kernel<ROWS>,
                                                                     kCompThreads, smem));
        if (occupancy < 1) occupancy = 1;
    }
    p.tiles = ceil_div(p.size, TILE);
    const size_t state_bytes = (size_t) p.tiles * 8;
    uint8_t *mem = (uint8_t *) scratch.device(256 + state_bytes);
    p.ticket = (uint32_t *) mem;
    p.state = (uint64_t *) (mem + 256);
    DJB_CUDA_CHECK(cudaMemsetAsync(mem, 0, 256 + state_bytes, stream));

    const uint32_t grid = std::min(p.tiles, dev.sm_count * (uint32_t) occupancy);
    compress_kernel<ROWS><<<grid, kCompThreads, smem, stream>>>(p);
    DJB_POST_LAUNCH();
}

uint32_t compress(cudaStream_t stream, const uint8_t *in, uint32_t size, uint32_t index_base,
                  uint32_t *out, uint32_t *count_dev, bool sync) {
    if (size == 0) { // cuda_ts.cpp:685-686
        if (count_dev)
            DJB_CUDA_CHECK(cudaMemsetAsync(count_dev, 0, sizeof(uint32_t), stream));
        return 0;
    }

    Scratch scratch(stream);
    uint32_t *pinned = scratch.pinned_words();

    CompressParams p{};
    p.in = in; p.out = out; p.size = size; p.index_base = index_base;
    p.vec = ((uintptr_t) in % 16) == 0;
    if (count_dev) {
        p.count_out = count_dev;
    } else {
        DJB_CUDA_CHECK(cudaHostGetDevicePointer((void **) &p.count_out, pinned, 0));
    }

    // small tiles keep all SMs busy on small masks (cuda_ts.cpp:693 draws its line at 4096)
    const DeviceProps &dev = device_props();
    if ((uint64_t) size >= (uint64_t) kCompThreads * kCompRowsBig * kCompUnit * dev.sm_count * 4)
        launch_compress<kCompRowsBig>(stream, p, scratch);
    else
        launch_compress<kCompRowsSmall>(stream, p, scratch);

    if (!sync)
        return 0;
    // The reference synchronises here as well (jitc_sync_thread, cuda_ts.cpp:759)
    if (count_dev) {
        DJB_CUDA_CHECK(cudaMemcpyAsync(pinned, count_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    }
    DJB_CUDA_CHECK(cudaStreamSynchronize(stream));
    return pinned[0];
}

} // namespace djb
