/*
 * prefix_small.cu -- unsegmented prefix reductions of small arrays (up to 128 KiB; u8 32, f16 64) in ONE CTA of 1024
 * threads with every load in flight at once.
 *
 * Wavefront-sized arrays (CUDAThreadState::block_prefix_reduce,
 * ext/drjit-core/src/cuda_ts.cpp:530-681, is called by the thousand on arrays of 2^10 .. 2^16 elements
 * in Mitsuba's wavefront loops). The tile kernel (scan_kernel.cuh) walked such an array with one CTA of
 * 256 threads tile by tile -- the load latency of every 32 KiB tile in series: 8.3 us per call at 2^14
 * u32 against 6.2 us for the reference's 16 small CTAs (profiles/r5i_small_sizes.txt). Here a thread
 * owns up to eight 128-bit units in a warp-striped arrangement (unit = row * 1024 + tid), issues all of
 * its loads before it touches any of them, scans its units in registers, the rows across the lanes with
 * shuffles, and the 8 x 32 (row, warp) totals with one more two-level scan in shared memory: two
 * barriers, no scratch, no descriptors, one launch. `carry_in` / `total_out` as in the tile kernel
 * (shards of a distributed scan).
 */
#include "common.cuh"
#include "runtime.h"

namespace djb {

constexpr uint32_t kSmallThreads = 1024, kSmallWarps = 32, kSmallMaxRows = 8;
/// Units per thread: at most 32 accumulators (u8: 2 units of 16 elements, f16: 4 of 8, 4- and 8-byte types: 8)
template <typename T> struct SmallRows {
    static constexpr uint32_t value = (16 / sizeof(T)) * kSmallMaxRows <= 32 ? kSmallMaxRows : 32 / (16 / sizeof(T));
};

template <typename T, typename Op>
__global__ void __launch_bounds__(kSmallThreads, 1)
prefix_small_kernel(const T *in, T *out, uint32_t size, uint32_t rows, uint32_t exclusive, uint32_t reverse,
                    const T *carry_in, T *total_out) {
    using A = acc_t<T>;
    constexpr uint32_t V = 16 / sizeof(T), kSmallRows = SmallRows<T>::value;
    __shared__ A wtot[kSmallRows * kSmallWarps];    // (row, warp) totals in scan order, then their exclusive scan
    __shared__ A wsum[kSmallRows];
    const A ident = Op::template identity<A>();
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t n_units = (size + V - 1) / V;

    // ---- all loads first. Scan-order unit s = row * 1024 + tid; reverse scans mirror units and elements
    //      (the host only sends reverse scans whose size is a multiple of V: the mirror image is aligned)
    Vec16<T> raw[kSmallRows];
    #pragma unroll
    for (uint32_t k = 0; k < kSmallRows; ++k) {
        const uint32_t s = k * kSmallThreads + tid;
        if (k < rows && s < n_units) {
            const uint32_t u = reverse ? n_units - 1 - s : s;
            if ((u + 1) * V <= size) {
                raw[k] = ld_vec<T>(in + (size_t) u * V);            // (coherent: `out` may alias `in`)
            } else {                                                // ragged last unit (forward scans only)
                #pragma unroll
                for (uint32_t e = 0; e < V; ++e)
                    raw[k].v[e] = u * V + e < size ? in[(size_t) u * V + e] : T();
            }
        }
    }

    // ---- unit-local scans, row scans over the lanes ---------------------------------------------
    A incl[kSmallRows][V];
    A unit_ex[kSmallRows];      // reduction of the units before mine in my warp's part of the row
    #pragma unroll
    for (uint32_t k = 0; k < kSmallRows; ++k) {
        if (k >= rows) break;                                   // (uniform)
        const uint32_t s = k * kSmallThreads + tid;
        A run = ident;
        #pragma unroll
        for (uint32_t e = 0; e < V; ++e) {
            const uint32_t ee = reverse ? V - 1 - e : e;
            const uint32_t pos = s * V + e;                     // position in scan order
            const A x = pos < size ? to_acc<A>(raw[k].v[ee]) : ident;
            run = Op::template apply<A>(run, x);
            incl[k][e] = run;
        }
        A v = run;
        #pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const A t = shfl_up(v, d);
            if (lane >= d) v = Op::template apply<A>(t, v);
        }
        A ex = shfl_up(v, 1);
        unit_ex[k] = lane == 0 ? ident : ex;
        if (lane == 31) wtot[k * kSmallWarps + warp] = v;
    }
    __syncthreads();

    // ---- exclusive scan of the rows x 32 warp totals (scan order = index order) by the first `rows` warps
    if (warp < rows) {
        const A t = wtot[tid];
        A v = t;
        #pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const A u = shfl_up(v, d);
            if (lane >= d) v = Op::template apply<A>(u, v);
        }
        if (lane == 31) wsum[warp] = v;
        A ex = shfl_up(v, 1);
        wtot[tid] = lane == 0 ? ident : ex;                     // exclusive inside the row
    }
    __syncthreads();
    A carry = carry_in ? to_acc<A>(*carry_in) : ident;

    // ---- combine and store ---------------------------------------------------------------------
    A row_base = carry;
    #pragma unroll
    for (uint32_t k = 0; k < kSmallRows; ++k) {
        if (k >= rows) break;
        const uint32_t s = k * kSmallThreads + tid;
        const A enter = Op::template apply<A>(Op::template apply<A>(row_base, wtot[k * kSmallWarps + warp]), unit_ex[k]);
        row_base = Op::template apply<A>(row_base, wsum[k]);
        if (s >= n_units)
            continue;
        Vec16<T> o;
        #pragma unroll
        for (uint32_t e = 0; e < V; ++e) {
            const A r = exclusive ? (e == 0 ? enter : Op::template apply<A>(enter, incl[k][e ? e - 1 : 0]))
                                  : Op::template apply<A>(enter, incl[k][e]);
            o.v[reverse ? V - 1 - e : e] = from_acc<T>(r);
        }
        const uint32_t u = reverse ? n_units - 1 - s : s;
        if ((u + 1) * V <= size) {
            *reinterpret_cast<uint4 *>(out + (size_t) u * V) = *reinterpret_cast<const uint4 *>(&o);
        } else {
            #pragma unroll
            for (uint32_t e = 0; e < V; ++e)
                if (u * V + e < size) out[(size_t) u * V + e] = o.v[e];
        }
    }
    if (total_out && tid == 0)
        *total_out = from_acc<T>(row_base);                     // carry (+) everything
}

template <typename T, typename Op>
static bool small_launch(cudaStream_t stream, uint32_t size, bool exclusive, bool reverse, const void *in, void *out,
                         const void *carry_in, void *total_out) {
    constexpr uint32_t V = 16 / sizeof(T);
    if (reverse && size % V != 0)
        return false;                   // (mirrored vectors need the array end on a vector boundary)
    const uint32_t n_units = ceil_div(size, V), rows = ceil_div(n_units, kSmallThreads);
    if (rows > SmallRows<T>::value)
        return false;
    prefix_small_kernel<T, Op><<<1, kSmallThreads, 0, stream>>>((const T *) in, (T *) out, size, rows, exclusive ? 1u : 0u,
                                                                reverse ? 1u : 0u, (const T *) carry_in, (T *) total_out);
    DJB_POST_LAUNCH();
    return true;
}

template <typename T> static bool small_ops_int(cudaStream_t s, int op, uint32_t size, bool ex, bool rev, const void *in,
                                                void *out, const void *ci, void *to) {
    switch (op) {
        case DRJIT_B200_OP_ADD: return small_launch<T, OpAdd>(s, size, ex, rev, in, out, ci, to);
        case DRJIT_B200_OP_MUL: return small_launch<T, OpMul>(s, size, ex, rev, in, out, ci, to);
        case DRJIT_B200_OP_MIN: return small_launch<T, OpMin>(s, size, ex, rev, in, out, ci, to);
        case DRJIT_B200_OP_MAX: return small_launch<T, OpMax>(s, size, ex, rev, in, out, ci, to);
        case DRJIT_B200_OP_AND: return small_launch<T, OpAnd>(s, size, ex, rev, in, out, ci, to);
        case DRJIT_B200_OP_OR:  return small_launch<T, OpOr>(s, size, ex, rev, in, out, ci, to);
        default: return false;
    }
}
template <typename T> static bool small_ops_minmax(cudaStream_t s, int op, uint32_t size, bool ex, bool rev, const void *in,
                                                   void *out, const void *ci, void *to) {
    if (op == DRJIT_B200_OP_MIN) return small_launch<T, OpMin>(s, size, ex, rev, in, out, ci, to);
    if (op == DRJIT_B200_OP_MAX) return small_launch<T, OpMax>(s, size, ex, rev, in, out, ci, to);
    return false;
}
template <typename T> static bool small_ops_float(cudaStream_t s, int op, uint32_t size, bool ex, bool rev, const void *in,
                                                  void *out, const void *ci, void *to) {
    switch (op) {
        case DRJIT_B200_OP_ADD: return small_launch<T, OpAdd>(s, size, ex, rev, in, out, ci, to);
        case DRJIT_B200_OP_MUL: return small_launch<T, OpMul>(s, size, ex, rev, in, out, ci, to);
        case DRJIT_B200_OP_MIN: return small_launch<T, OpMin>(s, size, ex, rev, in, out, ci, to);
        case DRJIT_B200_OP_MAX: return small_launch<T, OpMax>(s, size, ex, rev, in, out, ci, to);
        default: return false;
    }
}

/// Unsegmented prefix reduction of an array of at most 128 KiB (both pointers 16-byte aligned); false if
/// the configuration is left to the tile kernel
bool prefix_small(cudaStream_t stream, int vt, int op, uint32_t size, bool exclusive, bool reverse, const void *in,
                  void *out, const void *carry_in, void *total_out) {
    const uint32_t tsize = type_size(vt);
    if (tsize == 0 || size == 0 || (uint64_t) size * tsize > kSmallThreads * kSmallMaxRows * 16 ||
        ((uintptr_t) in % 16) || ((uintptr_t) out % 16))
        return false;
    const bool sign_agnostic = op == DRJIT_B200_OP_ADD || op == DRJIT_B200_OP_MUL ||
                               op == DRJIT_B200_OP_AND || op == DRJIT_B200_OP_OR;
    switch (vt) {
        case DRJIT_B200_VT_BOOL:
        case DRJIT_B200_VT_UINT8:  return small_ops_int<uint8_t>(stream, op, size, exclusive, reverse, in, out, carry_in, total_out);
        case DRJIT_B200_VT_UINT32: return small_ops_int<uint32_t>(stream, op, size, exclusive, reverse, in, out, carry_in, total_out);
        case DRJIT_B200_VT_UINT64: return small_ops_int<uint64_t>(stream, op, size, exclusive, reverse, in, out, carry_in, total_out);
        case DRJIT_B200_VT_INT32:
            return sign_agnostic ? small_ops_int<uint32_t>(stream, op, size, exclusive, reverse, in, out, carry_in, total_out)
                                 : small_ops_minmax<int32_t>(stream, op, size, exclusive, reverse, in, out, carry_in, total_out);
        case DRJIT_B200_VT_INT64:
            return sign_agnostic ? small_ops_int<uint64_t>(stream, op, size, exclusive, reverse, in, out, carry_in, total_out)
                                 : small_ops_minmax<int64_t>(stream, op, size, exclusive, reverse, in, out, carry_in, total_out);
        case DRJIT_B200_VT_FLOAT16: return small_ops_float<__half>(stream, op, size, exclusive, reverse, in, out, carry_in, total_out);
        case DRJIT_B200_VT_FLOAT32: return small_ops_float<float>(stream, op, size, exclusive, reverse, in, out, carry_in, total_out);
        case DRJIT_B200_VT_FLOAT64: return small_ops_float<double>(stream, op, size, exclusive, reverse, in, out, carry_in, total_out);
        default: return false;
    }
}

} // namespace djb
