/*
 * sort.cu -- dr.sort / dr.argsort: stable LSD radix sort whose passes carry the keys and the index
 * payload themselves.
 *
 * Reference: drjit/__init__.py:1698-1772 (`_radix_sort`). On the GPU it runs 8-bit passes, each made
 * of a JIT kernel that extracts the digit, jit_block_mkperm on the digit array (the reference's six
 * launches, ext/drjit-core/src/cuda_ts.cpp:788-975) and one random gather per carried array
 * (ordinal, index) through the resulting permutation: about 44 bytes of DRAM traffic per element and
 * pass, most of it in 32-byte sectors fetched for 4 useful bytes.
 *
 * Here a pass is: one histogram launch (keys read once, 4 B/element), one tiny offsets launch, and
 * one scatter launch that reads the keys (+ payload) of a tile, ranks them in input order and writes
 * keys and payload to their final places run by run -- 20 B/element and pass for key + index,
 * 12 B for keys only, no permutation array, no gathers. The order-preserving ordinal transform of
 * `_to_ordinal_32/_64` (drjit/__init__.py:1483-1520) is applied on the fly when the digit is
 * extracted; keys travel in their original representation.
 *
 * Ranking is the stable scheme of mkperm.cu (mkperm_tile_scatter_stable_kernel): every warp owns a
 * contiguous segment of the tile and walks it 32 consecutive keys at a time; lanes with equal digits
 * are found with one ballot per digit bit. Order inside a bucket = tile order (per-tile offsets of
 * the histogram pass) > warp order > step order > lane order = input order.
 */
#include "common.cuh"
#include "runtime.h"
#include "tma.cuh"

#include <algorithm>
#include <atomic>

namespace djb {

constexpr uint32_t kSortBits = 8, kSortBuckets = 1u << kSortBits;
constexpr uint32_t kSortThreads = 512, kSortKpt = 16, kSortTile = kSortThreads * kSortKpt;
constexpr uint32_t kSortWarps = kSortThreads / 32, kSortSeg = kSortTile / kSortWarps;

enum : uint32_t { kXformUnsigned = 0, kXformSigned = 1, kXformFloat = 2, kXformDescending = 4 };

struct SortParams {
    const void *keys_in;
    void *keys_out;
    const uint32_t *pay_in;     // nullptr: the payload of element i is i (first pass of an argsort)
    uint32_t *pay_out;          // nullptr: keys only
    uint32_t *tile_off;         // [tiles][256] keys of the bucket in earlier tiles of the same chunk
    uint32_t *rows;             // [chunks][256] chunk totals -> exclusive chunk offsets
    uint32_t *bucket_start;     // [256]
    uint32_t size, tiles, tiles_per_chunk, chunks, shift, xform;
};

/// Digit of a key: bits [shift, shift + 8) of its order-preserving unsigned image
/// (_to_ordinal_32 / _to_ordinal_64, drjit/__init__.py:1483-1520; `~` for descending order)
template <typename K> __device__ __forceinline__ uint32_t sort_digit(K k, uint32_t shift, uint32_t xform) {
    constexpr K top = (K) 1 << (sizeof(K) * 8 - 1);
    if (xform & kXformSigned) k ^= top;
    if (xform & kXformFloat)  k ^= ((K) 0 - (k >> (sizeof(K) * 8 - 1))) | top;
    if (xform & kXformDescending) k = ~k;
    return (uint32_t) (k >> shift) & (kSortBuckets - 1);
}

// ---------------------------------------------------------------------------
//  Pass kernel 1: per-tile bucket offsets inside a chunk of tiles + chunk totals
// ---------------------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const SortParams p) {
    __shared__ uint32_t whist[kSortWarps][kSortBuckets];     // running per-warp counts of this chunk
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < kSortWarps * kSortBuckets; i += kSortThreads) (&whist[0][0])[i] = 0;
    const K *keys = reinterpret_cast<const K *>(p.keys_in);
    const uint32_t first = blockIdx.x * p.tiles_per_chunk, end = min(first + p.tiles_per_chunk, p.tiles);
    const bool vec = (((uintptr_t) keys) & 15u) == 0;
    constexpr uint32_t PER_VEC = 16 / sizeof(K), VECS = kSortKpt / PER_VEC;
    __syncthreads();
    for (uint32_t tile = first; tile < end; ++tile) {
        const uint64_t base = (uint64_t) tile * kSortTile;
        const uint32_t n_tile = (uint32_t) min((uint64_t) kSortTile, (uint64_t) p.size - base);
        K key[kSortKpt];
        const bool full = vec && n_tile == kSortTile;
        if (full) {
            #pragma unroll
            for (uint32_t v = 0; v < VECS; ++v) {
                const Vec16<K> t = ld_stream<K>(reinterpret_cast<const uint4 *>(keys + base) + v * kSortThreads + tid);
                #pragma unroll
                for (uint32_t e = 0; e < PER_VEC; ++e) key[v * PER_VEC + e] = t.v[e];
            }
            if (tid == 0 && tile + 1 < end && (uint64_t) (tile + 2) * kSortTile <= p.size)
                bulk_prefetch_l2(keys + base + kSortTile, kSortTile * sizeof(K));
        } else {
            #pragma unroll
            for (uint32_t k = 0; k < kSortKpt; ++k) {
                const uint32_t i = k * kSortThreads + tid;
                key[k] = i < n_tile ? keys[base + i] : (K) 0;
            }
        }
        // snapshot of the running counts = keys of every bucket in the earlier tiles of the chunk
        if (tid < kSortBuckets) {
            uint32_t s = 0;
            #pragma unroll
            for (uint32_t w = 0; w < kSortWarps; ++w) s += whist[w][tid];
            p.tile_off[(size_t) tile * kSortBuckets + tid] = s;
        }
        __syncthreads();
        #pragma unroll
        for (uint32_t k = 0; k < kSortKpt; ++k) {
            const uint32_t i = full ? k : k * kSortThreads + tid;     // (full tiles: every slot holds a key)
            if (full || i < n_tile)
                atomicAdd(&whist[warp][sort_digit(key[k], p.shift, p.xform)], 1u);
        }
        __syncthreads();
    }
    if (tid < kSortBuckets) {
        uint32_t s = 0;
        #pragma unroll
        for (uint32_t w = 0; w < kSortWarps; ++w) s += whist[w][tid];
        p.rows[(size_t) blockIdx.x * kSortBuckets + tid] = s;
    }
}

// ---------------------------------------------------------------------------
//  Pass kernel 2 (one CTA): chunk totals -> exclusive chunk offsets, bucket totals -> bucket starts
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
sort_offsets_kernel(const SortParams p) {
    __shared__ uint32_t seg_sum[4][kSortBuckets];
    __shared__ uint32_t warp_sum[8];
    const uint32_t tid = threadIdx.x, seg = tid >> 8, b = tid & 255u, lane = tid & 31u;
    const uint32_t per = (p.chunks + 3) / 4, r0 = min(seg * per, p.chunks), r1 = min(r0 + per, p.chunks);
    uint32_t *col = p.rows + b;
    uint32_t sum = 0;
    {
        uint32_t r = r0;
        for (; r + 8 <= r1; r += 8) {
            uint32_t t[8];
            #pragma unroll
            for (uint32_t u = 0; u < 8; ++u) t[u] = col[(size_t) (r + u) * kSortBuckets];
            #pragma unroll
            for (uint32_t u = 0; u < 8; ++u) sum += t[u];
        }
        for (; r < r1; ++r) sum += col[(size_t) r * kSortBuckets];
    }
    seg_sum[seg][b] = sum;
    __syncthreads();
    uint32_t running = 0, total = 0;
    #pragma unroll
    for (uint32_t s = 0; s < 4; ++s) {
        if (s == seg) running = total;
        total += seg_sum[s][b];
    }
    {
        uint32_t r = r0;
        for (; r + 8 <= r1; r += 8) {
            uint32_t t[8];
            #pragma unroll
            for (uint32_t u = 0; u < 8; ++u) t[u] = col[(size_t) (r + u) * kSortBuckets];
            #pragma unroll
            for (uint32_t u = 0; u < 8; ++u) { col[(size_t) (r + u) * kSortBuckets] = running; running += t[u]; }
        }
        for (; r < r1; ++r) {
            const uint32_t t = col[(size_t) r * kSortBuckets];
            col[(size_t) r * kSortBuckets] = running;
            running += t;
        }
    }
    // exclusive scan of the 256 bucket totals (threads of segment 0 = warps 0..7)
    uint32_t incl = total;
    if (seg == 0) {
        #pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t t = shfl_up(incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) warp_sum[tid >> 5] = incl;
    }
    __syncthreads();
    if (seg == 0) {
        uint32_t wbase = 0;
        #pragma unroll
        for (uint32_t w = 0; w < 8; ++w)
            if (w < (tid >> 5)) wbase += warp_sum[w];
        p.bucket_start[b] = wbase + incl - total;
    }
}

// ---------------------------------------------------------------------------
//  Pass kernel 3: stable scatter of keys (+ payload)
// ---------------------------------------------------------------------------
/// peers &= (lanes whose digit agrees with mine in the bit `mask`)
__device__ __forceinline__ uint32_t sort_match_bit(uint32_t peers, uint32_t digit, uint32_t mask) {
    const bool one = digit & mask;
    const uint32_t v = __ballot_sync(kFullMask, one), kb = one ? 0xffffffffu : 0u;
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x90;" : "=r"(d) : "r"(peers), "r"(v), "r"(kb));    // peers & ~(v ^ kb)
    return d;
}

template <typename K, bool PAYLOAD>
constexpr uint32_t sort_scatter_smem() {
    return kSortWarps * kSortBuckets * 4 + kSortBuckets * 4 + kSortTile * (uint32_t) sizeof(K) + (PAYLOAD ? kSortTile * 4 : 0);
}

template <typename K, bool PAYLOAD>
__global__ void __launch_bounds__(kSortThreads, 2)
sort_scatter_kernel(const SortParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *whist = smem;                                   // [WARPS][256] per-warp counts -> running tile-local positions
    uint32_t *delta = smem + kSortWarps * kSortBuckets;       // [256] final position of the bucket's run minus its local start
    K *skey = reinterpret_cast<K *>(delta + kSortBuckets);    // [TILE] keys in bucket order
    uint32_t *spay = reinterpret_cast<uint32_t *>(skey + kSortTile);  // [TILE] payload in bucket order
    __shared__ uint32_t warp_sum[8];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    uint32_t *mine = whist + warp * kSortBuckets;
    const K *keys = reinterpret_cast<const K *>(p.keys_in);
    K *keys_out = reinterpret_cast<K *>(p.keys_out);

    for (uint32_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const uint64_t base = (uint64_t) tile * kSortTile;
        const uint32_t n_tile = (uint32_t) min((uint64_t) kSortTile, (uint64_t) p.size - base);

        // keys of step s of this warp's segment: seg + s * 32 + lane (coalesced, input order)
        K key[kSortKpt];
        #pragma unroll
        for (uint32_t s = 0; s < kSortKpt; ++s) {
            const uint32_t i = warp * kSortSeg + s * 32 + lane;
            key[s] = i < n_tile ? __ldg(keys + base + i) : (K) 0;
        }
        if (tid == 0) {
            const uint64_t next = (uint64_t) tile + gridDim.x;
            if (next < p.tiles) {
                if ((((uintptr_t) keys) & 15u) == 0 && (next + 1) * kSortTile <= p.size)
                    bulk_prefetch_l2(keys + next * kSortTile, kSortTile * sizeof(K));
                if (PAYLOAD && p.pay_in && (((uintptr_t) p.pay_in) & 15u) == 0 && (next + 1) * kSortTile <= p.size)
                    bulk_prefetch_l2(p.pay_in + next * kSortTile, kSortTile * 4);
                bulk_prefetch_l2(p.tile_off + next * kSortBuckets, kSortBuckets * 4);
            }
        }

        // ---- (1) per-warp histograms of the segments -------------------------------------------
        for (uint32_t b = lane; b < kSortBuckets; b += 32) mine[b] = 0;
        __syncwarp();
        #pragma unroll
        for (uint32_t s = 0; s < kSortKpt; ++s)
            if (warp * kSortSeg + s * 32 + lane < n_tile)
                atomicAdd(mine + sort_digit(key[s], p.shift, p.xform), 1u);
        __syncthreads();

        // ---- (2) prefix over the warps of each bucket, then over the buckets (threads 0..255) -----
        {
            uint32_t tot = 0, goff = 0;
            if (tid < kSortBuckets) {
                const uint32_t chunk = tile / p.tiles_per_chunk;
                goff = __ldg(p.tile_off + (size_t) tile * kSortBuckets + tid) +
                       __ldg(p.rows + (size_t) chunk * kSortBuckets + tid) + __ldg(p.bucket_start + tid);
                #pragma unroll
                for (uint32_t w = 0; w < kSortWarps; ++w) {
                    const uint32_t c = whist[w * kSortBuckets + tid];
                    whist[w * kSortBuckets + tid] = tot;
                    tot += c;
                }
            }
            uint32_t incl = tot;
            if (tid < kSortBuckets) {
                #pragma unroll
                for (uint32_t d = 1; d < 32; d <<= 1) {
                    const uint32_t t = shfl_up(incl, d);
                    if (lane >= d) incl += t;
                }
                if (lane == 31) warp_sum[warp] = incl;
            }
            __syncthreads();
            if (tid < kSortBuckets) {
                uint32_t wbase = 0;
                #pragma unroll
                for (uint32_t w = 0; w < 8; ++w)
                    if (w < warp) wbase += warp_sum[w];
                const uint32_t start = wbase + incl - tot;      // tile-local start of bucket `tid`
                #pragma unroll
                for (uint32_t w = 0; w < kSortWarps; ++w) whist[w * kSortBuckets + tid] += start;
                delta[tid] = goff - start;
            }
            __syncthreads();
        }

        // ---- (3) ranking walk in input order ------------------------------------------------------
        const uint32_t idx0 = (uint32_t) base;
        auto walk = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            #pragma unroll
            for (uint32_t s = 0; s < kSortKpt; ++s) {
                const uint32_t local = warp * kSortSeg + s * 32 + lane, d = sort_digit(key[s], p.shift, p.xform);
                const bool valid = FULL || local < n_tile;
                uint32_t peers = FULL ? kFullMask : __ballot_sync(kFullMask, valid);
                #pragma unroll
                for (uint32_t bit = 0; bit < kSortBits; ++bit)
                    peers = sort_match_bit(peers, d, 1u << bit);
                const uint32_t rank = __popc(peers & lanemask_lt());
                uint32_t pay = 0;
                if (PAYLOAD && valid) pay = p.pay_in ? __ldg(p.pay_in + base + local) : idx0 + local;
                uint32_t pos = 0;
                if (valid) pos = mine[d] + rank;
                __syncwarp();
                if (valid && rank == 0) mine[d] = pos + __popc(peers);  // lowest lane of the group
                __syncwarp();
                if (valid) {
                    skey[pos] = key[s];
                    if (PAYLOAD) spay[pos] = pay;
                }
            }
        };
        if (n_tile == kSortTile) walk(std::true_type{});
        else                     walk(std::false_type{});
        __syncthreads();

        // ---- (4) runs of equal digits are contiguous in shared memory and in the output ------------
        #pragma unroll 4
        for (uint32_t j = tid; j < n_tile; j += kSortThreads) {
            const K k = skey[j];
            const uint32_t at = delta[sort_digit(k, p.shift, p.xform)] + j;
            keys_out[at] = k;
            if (PAYLOAD) p.pay_out[at] = spay[j];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
//  Host side
// ---------------------------------------------------------------------------
template <typename K, bool PAYLOAD>
static void sort_pass(cudaStream_t stream, const SortParams &p, uint32_t grid) {
    constexpr uint32_t smem = sort_scatter_smem<K, PAYLOAD>();
    static std::atomic<bool> configured_on[kMaxDevices] = {};
    std::atomic<bool> &configured = configured_on[device_props().device % kMaxDevices];
    if (!configured.load(std::memory_order_acquire)) {
        DJB_CUDA_CHECK(cudaFuncSetAttribute(sort_scatter_kernel<K, PAYLOAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        configured.store(true, std::memory_order_release);
    }
    sort_hist_kernel<K><<<p.chunks, kSortThreads, 0, stream>>>(p);
    DJB_POST_LAUNCH();
    sort_offsets_kernel<<<1, 1024, 0, stream>>>(p);
    DJB_POST_LAUNCH();
    sort_scatter_kernel<K, PAYLOAD><<<grid, kSortThreads, smem, stream>>>(p);
    DJB_POST_LAUNCH();
}

template <typename K>
static void sort_impl(cudaStream_t stream, uint32_t xform, uint32_t size, const void *keys, void *keys_out,
                      uint32_t *index_out) {
    const DeviceProps &dev = device_props();
    const bool payload = index_out != nullptr;
    constexpr uint32_t passes = sizeof(K) * 8 / kSortBits;      // even: the last pass lands in the outputs

    SortParams p{};
    p.size = size; p.xform = xform;
    p.tiles = (uint32_t) ceil_div64(size, kSortTile);
    p.chunks = std::min(p.tiles, dev.sm_count * 2);
    p.tiles_per_chunk = ceil_div(p.tiles, p.chunks);
    p.chunks = ceil_div(p.tiles, p.tiles_per_chunk);
    const uint32_t grid = std::min(p.tiles, dev.sm_count * 2);

    Scratch scratch(stream);
    auto r256 = [](size_t v) { return (v + 255) & ~(size_t) 255; };
    const size_t key_bytes = (size_t) size * sizeof(K), pay_bytes = payload ? (size_t) size * 4 : 0,
                 off_bytes = (size_t) p.tiles * kSortBuckets * 4, rows_bytes = (size_t) p.chunks * kSortBuckets * 4;
    // ping-pong: in -> tmp -> out -> tmp -> out ...; without keys_out a second key buffer is needed
    scratch.reserve(r256(key_bytes) * (keys_out ? 1 : 2) + r256(pay_bytes) + r256(off_bytes) + r256(rows_bytes) + 1024 + 512);
    void *key_tmp = scratch.device(key_bytes);
    void *key_dst = keys_out ? keys_out : scratch.device(key_bytes);
    uint32_t *pay_tmp = payload ? (uint32_t *) scratch.device(pay_bytes) : nullptr;
    p.tile_off = (uint32_t *) scratch.device(off_bytes);
    p.rows = (uint32_t *) scratch.device(rows_bytes);
    p.bucket_start = (uint32_t *) scratch.device(kSortBuckets * 4);

    for (uint32_t pass = 0; pass < passes; ++pass) {
        const bool to_tmp = (pass & 1u) == 0;
        p.shift = pass * kSortBits;
        p.keys_in = pass == 0 ? keys : (to_tmp ? key_dst : key_tmp);
        p.keys_out = to_tmp ? key_tmp : key_dst;
        p.pay_in = pass == 0 ? nullptr : (to_tmp ? index_out : pay_tmp);
        p.pay_out = payload ? (to_tmp ? pay_tmp : index_out) : nullptr;
        if (payload) sort_pass<K, true>(stream, p, grid);
        else         sort_pass<K, false>(stream, p, grid);
    }
}

/// dr.sort / dr.argsort of one array (block_size == size). keys_out and/or index_out may be null.
void sort(cudaStream_t stream, int vt, uint32_t size, bool descending, const void *keys, void *keys_out,
          uint32_t *index_out) {
    if (size == 0 || (!keys_out && !index_out))
        return;
    uint32_t xform = descending ? kXformDescending : 0u;
    switch (vt) {
        case DRJIT_B200_VT_UINT32: sort_impl<uint32_t>(stream, xform | kXformUnsigned, size, keys, keys_out, index_out); break;
        case DRJIT_B200_VT_INT32:  sort_impl<uint32_t>(stream, xform | kXformSigned, size, keys, keys_out, index_out); break;
        case DRJIT_B200_VT_FLOAT32: sort_impl<uint32_t>(stream, xform | kXformFloat, size, keys, keys_out, index_out); break;
        case DRJIT_B200_VT_UINT64: sort_impl<uint64_t>(stream, xform | kXformUnsigned, size, keys, keys_out, index_out); break;
        case DRJIT_B200_VT_INT64:  sort_impl<uint64_t>(stream, xform | kXformSigned, size, keys, keys_out, index_out); break;
        case DRJIT_B200_VT_FLOAT64: sort_impl<uint64_t>(stream, xform | kXformFloat, size, keys, keys_out, index_out); break;
        default:
            raise(DRJIT_B200_EUNSUPPORTED, "drjit_b200_sort(): unsupported type %s!", type_name(vt));
    }
}

} // namespace djb
