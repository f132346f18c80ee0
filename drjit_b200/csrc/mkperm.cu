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
#include "common.cuh"
#include "runtime.h"

#include <cstdlib>
#include <cstring>

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
                    for (int e = 0; e < 4; ++e) atomicAdd(hist + t[u].v[e], 1u);
                }
        }
        for (uint64_t i = start + nvec * 4 + lane; i < end; i += 32)
            atomicAdd(hist + p.values[i], 1u);
    } else {
        for (uint64_t i = start + lane; i < end; i += 32)
            atomicAdd(hist + __ldg(p.values + i), 1u);
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
    const size_t B = p.bucket_count;
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
        p.totals[(size_t) group * B + b] = total;
}

// ---------------------------------------------------------------------------
//  Phase 3: per group, exclusive scan of bucket totals (+ table of non-empty buckets)
// ---------------------------------------------------------------------------
constexpr uint32_t kBucketScanThreads = 1024;

__global__ void __launch_bounds__(kBucketScanThreads)
mkperm_bucket_scan_kernel(const MkpermParams p, uint32_t *offsets, uint32_t *unique_out,
                          uint32_t *hist_out) {
    __shared__ uint32_t warp_sum[32], warp_uniq[32];
    __shared__ uint32_t carry_sum, carry_uniq;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t group = blockIdx.x, B = p.bucket_count;
    uint32_t *totals = p.totals + (size_t) group * B;

    if (tid == 0) { carry_sum = 0; carry_uniq = 0; }
    __syncthreads();

    for (uint32_t base = 0; base < B; base += kBucketScanThreads) {
        const uint32_t b = base + tid;
        const uint32_t n = b < B ? totals[b] : 0u, u = n != 0;
        if (hist_out && b < B) hist_out[b] = n;

        uint32_t vs = n, vu = u;
        #pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t ts = shfl_up(vs, d), tu = shfl_up(vu, d);
            if (lane >= d) { vs += ts; vu += tu; }
        }
        if (lane == 31) { warp_sum[warp] = vs; warp_uniq[warp] = vu; }
        __syncthreads();
        uint32_t ws = 0, wu = 0, ts_all = 0, tu_all = 0;
        #pragma unroll
        for (uint32_t w = 0; w < 32; ++w) {
            if (w == warp) { ws = ts_all; wu = tu_all; }
            ts_all += warp_sum[w]; tu_all += warp_uniq[w];
        }
        const uint32_t start = carry_sum + ws + vs - n,   // exclusive
                       slot = carry_uniq + wu + vu - u;
        if (b < B) {
            totals[b] = start;
            if (offsets && u) {       // quadruple layout: jit.h:2412-2419, mkperm.cuh:271-320
                uint4 q = make_uint4(b, start, n, 0u);
                *reinterpret_cast<uint4 *>(offsets + 4 * (size_t) slot) = q;
            }
        }
        __syncthreads();
        if (tid == 0) { carry_sum += ts_all; carry_uniq += tu_all; }
        __syncthreads();
    }
    if (tid == 0 && offsets) {
        offsets[4 * (size_t) B] = carry_uniq;  // cuda_ts.cpp:948-951
        if (unique_out) *unique_out = carry_uniq;
    }
}

// ---------------------------------------------------------------------------
//  Phase 4: scatter
// ---------------------------------------------------------------------------
template <MkpermMode Mode>
__global__ void mkperm_scatter_kernel(const MkpermParams p) {
    extern __shared__ uint32_t smem[];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u, nwarps = blockDim.x >> 5;
    const uint32_t group = blockIdx.x / p.ctas_per_group, cta = blockIdx.x - group * p.ctas_per_group;
    const uint32_t B = p.bucket_count;
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
                key[u] = ok[u] ? __ldg(p.values + i) : 0u;
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t active = __ballot_sync(kFullMask, ok[u]);
                if (active == 0) break;
                uint32_t pos = 0, peers = 0;
                if (ok[u]) {
                    peers = __match_any_sync(active, key[u]);
                    pos = ctr[key[u]] + __popc(peers & lanemask_lt());
                }
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
            uint32_t pos = atomicAdd(ctr + key, 1u);
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
//  Host side
// ---------------------------------------------------------------------------
static MkpermMode pick_mode(uint32_t bucket_count, uint32_t smem_budget, uint32_t &warps) {
    const uint64_t bytes = (uint64_t) bucket_count * 4;
    MkpermMode mode;
    uint32_t w = (uint32_t) std::min<uint64_t>(32, smem_budget / bytes);
    if (w >= 8) { mode = MkpermMode::Warp; warps = w; }
    else if (bytes <= smem_budget) { mode = MkpermMode::Cta; warps = 32; }
    else { mode = MkpermMode::Global; warps = 32; }

    // Developer override for A/B measurements: DRJIT_B200_MKPERM_MODE=warp|cta|global
    if (const char *env = getenv("DRJIT_B200_MKPERM_MODE")) {
        if (!strcmp(env, "cta") && bytes <= smem_budget) { mode = MkpermMode::Cta; warps = 32; }
        else if (!strcmp(env, "global")) { mode = MkpermMode::Global; warps = 32; }
        else if (!strcmp(env, "warp") && w >= 1) { mode = MkpermMode::Warp; warps = w; }
    }
    return mode;
}

template <MkpermMode Mode>
static void launch_phases(cudaStream_t stream, MkpermParams &p, uint32_t threads, uint32_t smem,
                          uint32_t *offsets_dev, uint32_t *unique_dev, uint32_t *hist_out,
                          cudaEvent_t table_ready) {
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
    mkperm_bucket_scan_kernel<<<p.n_groups, kBucketScanThreads, 0, stream>>>(p, offsets_dev, unique_dev, hist_out);
    DJB_POST_LAUNCH();
    if (table_ready)
        DJB_CUDA_CHECK(cudaEventRecord(table_ready, stream)); // cuda_ts.cpp:953 (before phase 4)
    mkperm_scatter_kernel<Mode><<<grid, threads, smem, stream>>>(p);
    DJB_POST_LAUNCH();
}

static cudaEvent_t mkperm_event() {
    static thread_local cudaEvent_t ev = nullptr;
    static thread_local int ev_device = -1;
    int device = device_props().device;
    if (!ev || ev_device != device) {
        DJB_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ev_device = device;
    }
    return ev;
}

/// Shared implementation. offsets: pinned host table (may be NULL). hist_out: optional device
/// array receiving the per-bucket counts (single group only). Returns the unique count when
/// `offsets` is given and there is a single group (after waiting for the table), else 0.
static uint32_t mkperm_impl(cudaStream_t stream, const uint32_t *values, uint32_t size,
                            uint32_t block_size, uint32_t bucket_count, uint32_t index_base,
                            uint32_t *perm, uint32_t *offsets, uint32_t *hist_out) {
    if (size == 0)
        return 0;
    if (bucket_count == 0) // cuda_ts.cpp:794-795 (jitc_fail)
        raise(DRJIT_B200_EFATAL, "jit_block_mkperm(): bucket_count cannot be zero!");
    if (block_size == 0 || block_size > size)
        raise(DRJIT_B200_EINVAL, "jit_block_mkperm(): invalid block size (size=%u, block_size=%u)!",
              size, block_size);

    const DeviceProps &dev = device_props();
    MkpermParams p{};
    p.values = values; p.perm = perm; p.size = size; p.block_size = block_size;
    p.bucket_count = bucket_count; p.index_base = index_base;
    p.n_groups = ceil_div(size, block_size);

    uint32_t warps = 32;
    const uint32_t smem_budget = dev.smem_optin - 1024;
    const MkpermMode mode = pick_mode(bucket_count, smem_budget, warps);
    // small sorting groups: do not spend more warps (= histogram rows) than the group can feed
    warps = std::max(1u, std::min(warps, ceil_div(block_size, 2048)));
    const uint32_t threads = warps * 32;

    // Rows: one per warp. Spread each group over enough CTAs to fill the machine once.
    const uint32_t min_row_elems = 2048;
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
    cudaEvent_t ev = want_table ? mkperm_event() : nullptr;
    switch (mode) {
        case MkpermMode::Warp:
            launch_phases<MkpermMode::Warp>(stream, p, threads, smem, offsets_dev, unique_dev, hist_out, ev);
            break;
        case MkpermMode::Cta:
            launch_phases<MkpermMode::Cta>(stream, p, threads, smem, offsets_dev, unique_dev, hist_out, ev);
            break;
        default:
            launch_phases<MkpermMode::Global>(stream, p, threads, 0, offsets_dev, unique_dev, hist_out, ev);
            break;
    }

    if (!want_table)
        return 0;
    DJB_CUDA_CHECK(cudaEventSynchronize(ev)); // cuda_ts.cpp:964-967: table valid, perm still in flight
    return pinned[1];
}

uint32_t block_mkperm(cudaStream_t stream, const uint32_t *values, uint32_t size, uint32_t block_size,
                      uint32_t bucket_count, uint32_t *perm, uint32_t *offsets) {
    return mkperm_impl(stream, values, size, block_size, bucket_count, 0, perm, offsets, nullptr);
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

} // namespace djb
