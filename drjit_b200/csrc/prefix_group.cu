/*
 * prefix_group.cu -- segmented prefix reductions of medium-sized blocks: every block is scanned by a
 * group of G lanes.
 *
 * dr.block_prefix_sum / cumsum along a trailing axis of a few dozen to a few thousand entries
 * (CUDAThreadState::block_prefix_reduce with block_size < size, ext/drjit-core/src/cuda_ts.cpp:530-681;
 * the reference runs the same look-back kernel as for a flat scan, resources/block_prefix_reduce.cuh:46-214).
 *
 * Blocks are independent, so nothing has to travel between CTAs: a group of G lanes (G = 4, 8, 16 or
 * 32, chosen from the block size) walks one block in chunks of G 128-bit units -- unit-local scan in
 * registers, one log2(G)-step shuffle scan, running carry in a register -- and a warp holds 32 / G
 * blocks at once. No tiles, no descriptors, no shared memory, no barriers, and none of the per-element
 * head arithmetic of the general segmented kernel (scan_kernel.cuh), which runs at 40-44 % of the copy
 * bandwidth whatever the block size (profiles/r5d_scanseg_sweep_general_kernel.txt).
 *
 * Units are aligned to the array base, not to the block: the first and last unit of a block may be
 * shared with its neighbours; foreign elements enter the scan as the identity and are not stored
 * (scalar stores at the ragged ends, STG.128 everywhere else), so in-place operation stays safe.
 * A lane keeps kDepth - 1 units in flight ahead of the one it scans (register ring, statically
 * indexed): with one unit of look-ahead the kernel sat on the load latency (ncu: 70 % of the stall
 * samples at the first use of the loaded unit, profiles/r5e_ncu_scanseg.md).
 */
#include "common.cuh"
#include "runtime.h"

#include <atomic>
#include <cstdlib>

namespace djb {

constexpr uint32_t kGroupBlocksMaxBytes = 1024 * 1024;
constexpr uint32_t kGroupBlocksFewBytes = 128 * 1024;  // longer blocks only with a warp on (nearly) every scheduler slot
constexpr uint32_t kGroupThreads = 256;
constexpr uint32_t kGroupDepth = 4;         // units in flight per lane (register ring)

template <typename T, typename Op, uint32_t G, bool REV>
__global__ void __launch_bounds__(kGroupThreads)
prefix_group_blocks_kernel(const T *in, T *out, uint32_t size, uint32_t bs, uint32_t n_blocks,
                           uint32_t iters, uint32_t exclusive) {
    using A = acc_t<T>;
    constexpr uint32_t V = 16 / sizeof(T), GPW = 32 / G, D = kGroupDepth;
    const A ident = Op::template identity<A>();
    const uint32_t lane = threadIdx.x & 31u, gl = lane & (G - 1u), sub = lane / G;
    const uint64_t warp = ((uint64_t) blockIdx.x * kGroupThreads + threadIdx.x) >> 5,
                   n_warps = (uint64_t) gridDim.x * (kGroupThreads / 32);

    // (all lanes of a warp run the same number of rounds and chunks: the shuffles are warp-wide)
    for (uint64_t b0 = warp * GPW; b0 < n_blocks; b0 += n_warps * GPW) {
        const uint64_t b = b0 + sub;
        const bool active = b < n_blocks;
        const uint64_t start = active ? b * bs : 0;
        const uint32_t len = active ? (uint32_t) (start + bs <= size ? (uint64_t) bs : (uint64_t) size - start) : 0u;
        // everything below is relative to the first element of the block's first unit (32-bit arithmetic):
        // the block is [lo, hi), its units are 0 .. n_units - 1
        const uint64_t ubase = start & ~(uint64_t) (V - 1);
        const uint32_t lo = (uint32_t) (start - ubase), hi = lo + len;
        const uint32_t n_units = active ? (hi + V - 1) / V : 0u;
        const T *bin = in + ubase;
        T *bout = out + ubase;
        const bool tail_partial = ubase + (uint64_t) n_units * V > size;     // the array ends inside the last unit

        // unit k of the block in scan order (mirrored for reverse scans)
        auto load = [&](uint32_t k, Vec16<T> &raw) {
            if (k >= n_units)
                return;
            const uint32_t u = REV ? n_units - 1u - k : k;
            if (!tail_partial || u + 1u < n_units) {
                raw = ld_vec<T>(bin + u * V);               // (coherent: `out` may alias `in`)
            } else {
                #pragma unroll
                for (uint32_t e = 0; e < V; ++e)
                    raw.v[e] = (u * V + e >= lo && u * V + e < hi) ? bin[u * V + e] : T();
            }
        };

        A carry = ident;
        Vec16<T> ring[D] = {};
        #pragma unroll
        for (uint32_t j = 0; j + 1 < D; ++j)
            if (j < iters) load(j * G + gl, ring[j]);

        for (uint32_t it0 = 0; it0 < iters; it0 += D) {
            #pragma unroll
            for (uint32_t j = 0; j < D; ++j) {
                const uint32_t it = it0 + j;
                if (it >= iters)                    // (warp-uniform)
                    break;
                const uint32_t k = it * G + gl;
                if (it + D - 1 < iters)
                    load(k + (D - 1) * G, ring[(j + D - 1) % D]);
                const Vec16<T> &cur = ring[j];
                const bool valid = k < n_units;
                const uint32_t rel = (REV ? n_units - 1u - k : k) * V;
                // bit e: element e of the unit belongs to the block
                uint32_t m = 0;
                if (valid) {
                    const uint32_t first = lo > rel ? lo - rel : 0u, last = hi - rel < V ? hi - rel : V;
                    m = ((1u << last) - 1u) & ~((1u << first) - 1u);
                }
                const bool full = m == (1u << V) - 1u;

                // unit-local scan in scan order; foreign elements are the identity
                A incl[V], excl[V];
                A run = ident;
                #pragma unroll
                for (uint32_t e = 0; e < V; ++e) {
                    const uint32_t ee = REV ? V - 1 - e : e;                // element of the unit
                    const A x = (full || ((m >> ee) & 1u)) ? to_acc<A>(cur.v[ee]) : ident;
                    excl[e] = run;
                    run = Op::template apply<A>(run, x);
                    incl[e] = run;
                }
                // scan of the unit totals over the group's lanes
                A v = run;
                #pragma unroll
                for (uint32_t d = 1; d < G; d <<= 1) {
                    const A t = __shfl_up_sync(kFullMask, v, d, G);
                    if (gl >= d) v = Op::template apply<A>(t, v);
                }
                A ex = __shfl_up_sync(kFullMask, v, 1, G);
                if (gl == 0) ex = ident;
                const A chunk_total = __shfl_sync(kFullMask, v, G - 1u, G);
                const A prefix = Op::template apply<A>(carry, ex);
                carry = Op::template apply<A>(carry, chunk_total);

                if (valid) {
                    Vec16<T> o;
                    #pragma unroll
                    for (uint32_t e = 0; e < V; ++e)
                        o.v[REV ? V - 1 - e : e] = from_acc<T>(Op::template apply<A>(prefix, exclusive ? excl[e] : incl[e]));
                    if (full) {
                        *reinterpret_cast<uint4 *>(bout + rel) = *reinterpret_cast<const uint4 *>(&o);
                    } else {
                        #pragma unroll
                        for (uint32_t e = 0; e < V; ++e)
                            if ((m >> e) & 1u)
                                bout[rel + e] = o.v[e];
                    }
                }
            }
        }
    }
}

template <typename T, typename Op, uint32_t G, bool REV>
static void launch_group(cudaStream_t stream, const T *in, T *out, uint32_t size, uint32_t bs, uint32_t n_blocks,
                         uint32_t iters, bool exclusive, uint64_t warps) {
    const DeviceProps &dev = device_props();
    auto kernel = prefix_group_blocks_kernel<T, Op, G, REV>;
    // one resident wave; the kernel strides over the blocks
    static std::atomic<int> occupancy_of[kMaxDevices] = {};
    int occupancy = occupancy_of[dev.device % kMaxDevices].load(std::memory_order_acquire);
    if (occupancy == 0) {                                   // (idempotent: a race only repeats the query)
        DJB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, kernel, kGroupThreads, 0));
        if (occupancy < 1) occupancy = 1;
        occupancy_of[dev.device % kMaxDevices].store(occupancy, std::memory_order_release);
    }
    const uint32_t grid = (uint32_t) std::min<uint64_t>(ceil_div64(warps, kGroupThreads / 32),
                                                        (uint64_t) dev.sm_count * (uint32_t) occupancy);
    kernel<<<grid, kGroupThreads, 0, stream>>>(in, out, size, bs, n_blocks, iters, exclusive ? 1u : 0u);
    DJB_POST_LAUNCH();
}

/// Lanes per block. A block that starts inside a unit touches one unit more than it fills; blocks whose
/// size is a multiple of the unit all start on unit boundaries.
template <typename T> static uint32_t group_units_max(uint32_t bs) {
    constexpr uint32_t V = 16 / sizeof(T);
    return ceil_div(bs, V) + (bs % V ? 1u : 0u);
}

/// Smallest share of lane slots carrying data at which the group kernel is used (see group_dispatch)
template <typename T> static double group_min_fill() {
#if defined(DRJIT_B200_EXPERIMENTS)
    static const char *env = getenv("DRJIT_B200_SCAN_GROUP_FILL");      // break-even sweeps
    if (env) return atof(env);
#endif
    // break-even against the general kernel (f32 43 %, u32 40 %, f64 30 %, f16 34 %, u8 19 % of the copy
    // bandwidth at any block size) with the group kernel forced on: one chunk per block runs at about
    // 0.65 x fill of the copy bandwidth for 4- and 8-byte types, 0.42 x fill for f16, 0.26 x fill for u8
    // (profiles/r5e_scanseg_sweep_group_forced.txt)
    return sizeof(T) == 8 ? 0.5 : sizeof(T) == 4 ? 0.67 : sizeof(T) == 2 ? 0.8 : 0.75;
}

template <typename T, typename Op>
static bool group_dispatch(cudaStream_t stream, uint32_t size, uint32_t bs, bool exclusive, bool reverse,
                           const void *in_, void *out_) {
    const uint32_t units_max = group_units_max<T>(bs);
    const uint32_t G = units_max <= 4 ? 4u : units_max <= 8 ? 8u : units_max <= 16 ? 16u : 32u;
    const uint32_t iters = ceil_div(units_max, G);
    const uint32_t n_blocks = ceil_div(size, bs);
    const uint64_t warps = ceil_div64(n_blocks, 32 / G);
    const DeviceProps &dev = device_props();
    // long blocks need enough of them to fill the machine: a warp walks its block chunk by chunk
    // (a warp has three 512-byte loads in flight: a quarter of the resident warps cannot fill the memory
    //  pipeline alone, but their blocks are short enough for the tail not to matter; blocks beyond 128 KiB
    //  need half of the 32 warps per SM that fit)
    const uint32_t min_warps_per_sm = (uint64_t) bs * sizeof(T) > kGroupBlocksFewBytes ? 16u : 4u;
    if (iters > 4 && warps < (uint64_t) dev.sm_count * min_warps_per_sm)
        return false;
    // Lane slots that carry data. Short blocks leave lanes idle (a block of 9 f32 fills 2.25 of its 4
    // lanes' units) while the general kernel's cost per element does not depend on the block size: below
    // the measured break-even the call is left to it.
    const double fill = (double) bs * sizeof(T) / 16.0 / ((double) iters * G);
    if (fill < group_min_fill<T>())
        return false;
    const T *in = (const T *) in_; T *out = (T *) out_;
    #define DJB_GROUP(G_) \
        (reverse ? launch_group<T, Op, G_, true>(stream, in, out, size, bs, n_blocks, iters, exclusive, warps) \
                 : launch_group<T, Op, G_, false>(stream, in, out, size, bs, n_blocks, iters, exclusive, warps))
    switch (G) {
        case 4: DJB_GROUP(4); break;
        case 8: DJB_GROUP(8); break;
        case 16: DJB_GROUP(16); break;
        default: DJB_GROUP(32); break;
    }
    #undef DJB_GROUP
    return true;
}

template <typename T> static bool group_ops_int(cudaStream_t s, int op, uint32_t size, uint32_t bs, bool ex, bool rev,
                                                const void *in, void *out) {
    switch (op) {
        case DRJIT_B200_OP_ADD: return group_dispatch<T, OpAdd>(s, size, bs, ex, rev, in, out);
        case DRJIT_B200_OP_MUL: return group_dispatch<T, OpMul>(s, size, bs, ex, rev, in, out);
        case DRJIT_B200_OP_MIN: return group_dispatch<T, OpMin>(s, size, bs, ex, rev, in, out);
        case DRJIT_B200_OP_MAX: return group_dispatch<T, OpMax>(s, size, bs, ex, rev, in, out);
        case DRJIT_B200_OP_AND: return group_dispatch<T, OpAnd>(s, size, bs, ex, rev, in, out);
        case DRJIT_B200_OP_OR:  return group_dispatch<T, OpOr>(s, size, bs, ex, rev, in, out);
        default: return false;
    }
}
template <typename T> static bool group_ops_minmax(cudaStream_t s, int op, uint32_t size, uint32_t bs, bool ex, bool rev,
                                                   const void *in, void *out) {
    if (op == DRJIT_B200_OP_MIN) return group_dispatch<T, OpMin>(s, size, bs, ex, rev, in, out);
    if (op == DRJIT_B200_OP_MAX) return group_dispatch<T, OpMax>(s, size, bs, ex, rev, in, out);
    return false;
}
template <typename T> static bool group_ops_float(cudaStream_t s, int op, uint32_t size, uint32_t bs, bool ex, bool rev,
                                                  const void *in, void *out) {
    switch (op) {
        case DRJIT_B200_OP_ADD: return group_dispatch<T, OpAdd>(s, size, bs, ex, rev, in, out);
        case DRJIT_B200_OP_MUL: return group_dispatch<T, OpMul>(s, size, bs, ex, rev, in, out);
        case DRJIT_B200_OP_MIN: return group_dispatch<T, OpMin>(s, size, bs, ex, rev, in, out);
        case DRJIT_B200_OP_MAX: return group_dispatch<T, OpMax>(s, size, bs, ex, rev, in, out);
        default: return false;      // (the caller raises the reference's error for And / Or on floats)
    }
}

bool prefix_group_blocks(cudaStream_t stream, int vt, int op, uint32_t size, uint32_t bs, bool exclusive,
                         bool reverse, const void *in, void *out) {
    const uint32_t tsize = type_size(vt);
    if (tsize == 0 || bs < 2 || bs >= size || (uint64_t) bs * tsize > kGroupBlocksMaxBytes ||
        ((uintptr_t) in % 16) || ((uintptr_t) out % 16))
        return false;
#if defined(DRJIT_B200_EXPERIMENTS)
    static const int off = getenv("DRJIT_B200_SCAN_NO_GROUP") ? atoi(getenv("DRJIT_B200_SCAN_NO_GROUP")) : 0;     // A/B
    if (off) return false;
#endif
    const bool sign_agnostic = op == DRJIT_B200_OP_ADD || op == DRJIT_B200_OP_MUL ||
                               op == DRJIT_B200_OP_AND || op == DRJIT_B200_OP_OR;
    switch (vt) {
        case DRJIT_B200_VT_BOOL:
        case DRJIT_B200_VT_UINT8:  return group_ops_int<uint8_t>(stream, op, size, bs, exclusive, reverse, in, out);
        case DRJIT_B200_VT_UINT32: return group_ops_int<uint32_t>(stream, op, size, bs, exclusive, reverse, in, out);
        case DRJIT_B200_VT_UINT64: return group_ops_int<uint64_t>(stream, op, size, bs, exclusive, reverse, in, out);
        case DRJIT_B200_VT_INT32:
            return sign_agnostic ? group_ops_int<uint32_t>(stream, op, size, bs, exclusive, reverse, in, out)
                                 : group_ops_minmax<int32_t>(stream, op, size, bs, exclusive, reverse, in, out);
        case DRJIT_B200_VT_INT64:
            return sign_agnostic ? group_ops_int<uint64_t>(stream, op, size, bs, exclusive, reverse, in, out)
                                 : group_ops_minmax<int64_t>(stream, op, size, bs, exclusive, reverse, in, out);
        case DRJIT_B200_VT_FLOAT16: return group_ops_float<__half>(stream, op, size, bs, exclusive, reverse, in, out);
        case DRJIT_B200_VT_FLOAT32: return group_ops_float<float>(stream, op, size, bs, exclusive, reverse, in, out);
        case DRJIT_B200_VT_FLOAT64: return group_ops_float<double>(stream, op, size, bs, exclusive, reverse, in, out);
        default: return false;
    }
}

} // namespace djb
