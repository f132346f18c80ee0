/*
 * scatter_reduce.cu -- hand-written scatter-reduce ("histogram") kernels.
 *
 * The reference has no precompiled kernel for dr.scatter_reduce: jitc_cuda_render_scatter_reduce
 * (ext/drjit-core/src/cuda_scatter.cpp:246-354) emits `red.global.<op>.<type>` into the fused
 * JIT kernel, optionally preceded by a warp-level pre-reduction (`ReduceMode::Local`,
 * cuda_scatter.cpp:125-244) built on match.any. This file provides the standalone primitive
 * for the benchmarked case (values + indices in memory -> bins):
 *   - DIRECT   : 128-bit loads of 4 values + 4 indices per thread, one fire-and-forget
 *                reduction (`red.global`, no return value) per element;
 *   - LOCAL    : the same, preceded by a match.any pre-reduction so that lanes hitting the
 *                same bin issue a single atomic (pays off for heavily contended bins);
 *   - PRIVATE  : bins small enough for shared memory (<= 55K 4-byte bins) are accumulated
 *                per CTA in shared memory and flushed once -- no L2 atomic traffic per element.
 * Float min/max use the signed/unsigned integer-atomic trick of cuda_scatter.cpp:74-106; f16
 * add/min/max use the two-wide f16 reductions with an identity partner (cuda_scatter.cpp:291-332).
 */
#include "common.cuh"
#include "atomic_ops.cuh"
#include "runtime.h"

namespace djb {

constexpr uint32_t kScThreads = 256;

struct ScatterParams {
    void *target;
    const void *value;
    const uint32_t *index;
    const uint8_t *mask;
    uint32_t size, target_size;
    uint8_t vec;
};

/// Warp pre-reduction: lanes with equal index combine, the lowest lane keeps the result
template <typename Op, typename T>
__device__ __forceinline__ bool warp_combine(uint32_t active, uint32_t idx, T &v) {
    const uint32_t peers = __match_any_sync(active, idx);
    const uint32_t lane = lane_id();
    const bool leader = (peers & lanemask_lt()) == 0;
    // (no early exit for lanes without peers: every lane of `active` must reach the
    // ballots / shuffles below)
    // walk the peer set: every lane pulls the values of its peers (leader's result is used)
    T acc = v;
    uint32_t rest = peers & ~(1u << lane);
    // all lanes of `active` must take part in the shuffles
    uint32_t pending = __ballot_sync(active, rest != 0);
    while (pending) {
        const uint32_t src = rest ? (uint32_t) __ffs(rest) - 1 : lane;
        const T other = __shfl_sync(active, v, src);
        if (rest) {
            acc = Op::template apply<T>(acc, other);
            rest &= rest - 1;
        }
        pending = __ballot_sync(active, rest != 0);
    }
    v = acc;
    return leader;
}

template <typename T, typename Op, bool LOCAL>
__global__ void __launch_bounds__(kScThreads)
scatter_reduce_kernel(const ScatterParams p) {
    T *target = (T *) p.target;
    const T *value = (const T *) p.value;
    const uint64_t gtid = (uint64_t) blockIdx.x * kScThreads + threadIdx.x,
                   gstride = (uint64_t) gridDim.x * kScThreads;

    auto emit = [&](bool ok, uint32_t idx, T v) {
        if constexpr (LOCAL) {
            const uint32_t active = __ballot_sync(kFullMask, ok);
            if (ok) {
                bool leader = true;
                if constexpr (!std::is_same<T, __half>::value)
                    leader = warp_combine<Op, T>(active, idx, v);
                if (leader) AtomicOp<Op, T>::apply(target + idx, v);
            }
        } else {
            if (ok) AtomicOp<Op, T>::apply(target + idx, v);
        }
    };

    if (p.vec) {
        // 4 indices per 128-bit load; values in matching 4-element groups
        constexpr uint32_t VB = 4 * sizeof(T);       // bytes of 4 values
        const uint64_t ngroups = p.size / 4;
        const uint64_t ngroups_rounded = LOCAL ? (ngroups + gstride - 1) / gstride * gstride : ngroups;
        for (uint64_t g = gtid; g < ngroups_rounded; g += gstride) {
            const bool in_range = g < ngroups;
            uint32_t idx[4] = { 0, 0, 0, 0 };
            T val[4];
            uint32_t m = 0x01010101u;
            if (in_range) {
                const Vec16<uint32_t> iv = ld_stream<uint32_t>(p.index + g * 4);
                #pragma unroll
                for (int e = 0; e < 4; ++e) idx[e] = iv.v[e];
                if constexpr (VB == 16) {
                    const Vec16<T> vv = ld_stream<T>(value + g * 4);
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) val[e] = vv.v[e];
                } else if constexpr (VB == 32) {
                    const Vec16<T> v0 = ld_stream<T>(value + g * 4), v1 = ld_stream<T>(value + g * 4 + 2);
                    val[0] = v0.v[0]; val[1] = v0.v[1]; val[2] = v1.v[0]; val[3] = v1.v[1];
                } else {
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) val[e] = value[g * 4 + e];
                }
                if (p.mask) m = *reinterpret_cast<const uint32_t *>(p.mask + g * 4);
            }
            #pragma unroll
            for (int e = 0; e < 4; ++e)
                emit(in_range && ((m >> (8 * e)) & 0xffu) != 0, idx[e], val[e]);
        }
        const uint64_t tail = ngroups * 4;
        if (LOCAL) {
            if (blockIdx.x == 0 && threadIdx.x < 32) {
                const uint64_t i = tail + threadIdx.x;
                const bool ok = i < p.size && (!p.mask || p.mask[i]);
                emit(ok, ok ? p.index[i] : 0u, ok ? value[i] : T());
            }
        } else if (gtid < p.size - tail) {
            const uint64_t i = tail + gtid;
            emit(!p.mask || p.mask[i], p.index[i], value[i]);
        }
    } else {
        const uint64_t n = p.size;
        const uint64_t n_rounded = LOCAL ? (n + gstride - 1) / gstride * gstride : n;
        for (uint64_t i = gtid; i < n_rounded; i += gstride) {
            const bool ok = i < n && (!p.mask || p.mask[i]);
            emit(ok, ok ? p.index[i] : 0u, ok ? value[i] : T());
        }
    }
}

/// Privatised bins in shared memory (Add on 4-byte types); one flush per CTA
template <typename T>
__global__ void __launch_bounds__(kScThreads)
scatter_add_private_kernel(const ScatterParams p) {
    extern __shared__ uint32_t smem_raw[];
    T *bins = reinterpret_cast<T *>(smem_raw);
    for (uint32_t i = threadIdx.x; i < p.target_size; i += kScThreads) bins[i] = T(0);
    __syncthreads();

    const T *value = (const T *) p.value;
    const uint64_t gtid = (uint64_t) blockIdx.x * kScThreads + threadIdx.x,
                   gstride = (uint64_t) gridDim.x * kScThreads;
    if (p.vec) {
        const uint64_t ngroups = p.size / 4;
        for (uint64_t g = gtid; g < ngroups; g += gstride) {
            const Vec16<uint32_t> iv = ld_stream<uint32_t>(p.index + g * 4);
            const Vec16<T> vv = ld_stream<T>(value + g * 4);
            const uint32_t m = p.mask ? *reinterpret_cast<const uint32_t *>(p.mask + g * 4) : 0x01010101u;
            #pragma unroll
            for (int e = 0; e < 4; ++e)
                if ((m >> (8 * e)) & 0xffu) atomicAdd(bins + iv.v[e], vv.v[e]);
        }
        const uint64_t tail = ngroups * 4;
        if (gtid < p.size - tail) {
            const uint64_t i = tail + gtid;
            if (!p.mask || p.mask[i]) atomicAdd(bins + p.index[i], value[i]);
        }
    } else {
        for (uint64_t i = gtid; i < p.size; i += gstride)
            if (!p.mask || p.mask[i]) atomicAdd(bins + p.index[i], value[i]);
    }
    __syncthreads();
    T *target = (T *) p.target;
    for (uint32_t i = threadIdx.x; i < p.target_size; i += kScThreads) {
        const T v = bins[i];
        if (v != T(0)) atomicAdd(target + i, v);
    }
}

template <typename T, typename Op>
static void launch_scatter(cudaStream_t stream, int mode, ScatterParams &p) {
    const DeviceProps &dev = device_props();
    p.vec = ((uintptr_t) p.value % 16) == 0 && ((uintptr_t) p.index % 16) == 0 &&
            (!p.mask || ((uintptr_t) p.mask % 4) == 0);

    const uint64_t per_cta = (uint64_t) kScThreads * 4 * 4; // 4 iterations of 4 elements per thread
    uint32_t grid = (uint32_t) std::min<uint64_t>(ceil_div64(p.size, per_cta), (uint64_t) dev.sm_count * 8 * 4);
    grid = std::max(grid, 1u);

    if constexpr (std::is_same<Op, OpAdd>::value && sizeof(T) == 4) {
        const uint64_t bytes = (uint64_t) p.target_size * sizeof(T);
        // private bins pay off once every CTA sees many more elements than bins
        if (mode != DRJIT_B200_MODE_DIRECT && bytes <= dev.smem_optin - 1024 &&
            (uint64_t) p.size >= 8ull * p.target_size) {
            uint32_t g = std::min<uint32_t>(dev.sm_count * (bytes <= 100 * 1024 ? 2 : 1),
                                            (uint32_t) std::max<uint64_t>(1, p.size / (4ull * p.target_size + 4096)));
            g = std::max(g, 1u);
            if (bytes > 48 * 1024)
                DJB_CUDA_CHECK(cudaFuncSetAttribute(scatter_add_private_kernel<T>,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
            scatter_add_private_kernel<T><<<g, kScThreads, bytes, stream>>>(p);
            DJB_POST_LAUNCH();
            return;
        }
    }
    if (mode == DRJIT_B200_MODE_LOCAL)
        scatter_reduce_kernel<T, Op, true><<<grid, kScThreads, 0, stream>>>(p);
    else
        scatter_reduce_kernel<T, Op, false><<<grid, kScThreads, 0, stream>>>(p);
    DJB_POST_LAUNCH();
}

void scatter_reduce(cudaStream_t stream, int vt, int op, int mode, void *target, uint32_t target_size,
                    const void *value, const uint32_t *index, const uint8_t *mask, uint32_t size) {
    if (size == 0)
        return;
    if (mode < DRJIT_B200_MODE_AUTO || mode > DRJIT_B200_MODE_PERMUTE)
        raise(DRJIT_B200_EINVAL, "jit_var_scatter(): invalid reduction mode!");
    ScatterParams p{};
    p.target = target; p.value = value; p.index = index; p.mask = mask;
    p.size = size; p.target_size = target_size;

    // capability table: src/op.cpp:2735-2822
    auto unsupported = [&]() {
        raise(DRJIT_B200_EUNSUPPORTED,
              "jit_var_scatter(): the CUDA backend does not support the requested type of atomic "
              "reduction (%s) for variables of type (%s)", op_name(op), type_name(vt));
    };
    const bool sign_agnostic = op == DRJIT_B200_OP_ADD || op == DRJIT_B200_OP_AND || op == DRJIT_B200_OP_OR;

#define DJB_SC(T, OP) launch_scatter<T, OP>(stream, mode, p)
    switch (vt) {
        case DRJIT_B200_VT_INT32:
        case DRJIT_B200_VT_UINT32:
            if (op == DRJIT_B200_OP_ADD) DJB_SC(uint32_t, OpAdd);
            else if (op == DRJIT_B200_OP_AND) DJB_SC(uint32_t, OpAnd);
            else if (op == DRJIT_B200_OP_OR) DJB_SC(uint32_t, OpOr);
            else if (op == DRJIT_B200_OP_MIN) { if (vt == DRJIT_B200_VT_INT32) DJB_SC(int32_t, OpMin); else DJB_SC(uint32_t, OpMin); }
            else if (op == DRJIT_B200_OP_MAX) { if (vt == DRJIT_B200_VT_INT32) DJB_SC(int32_t, OpMax); else DJB_SC(uint32_t, OpMax); }
            else unsupported();
            break;
        case DRJIT_B200_VT_INT64:
        case DRJIT_B200_VT_UINT64:
            if (op == DRJIT_B200_OP_ADD) DJB_SC(uint64_t, OpAdd);
            else if (op == DRJIT_B200_OP_AND) DJB_SC(uint64_t, OpAnd);
            else if (op == DRJIT_B200_OP_OR) DJB_SC(uint64_t, OpOr);
            else if (op == DRJIT_B200_OP_MIN) { if (vt == DRJIT_B200_VT_INT64) DJB_SC(int64_t, OpMin); else DJB_SC(uint64_t, OpMin); }
            else if (op == DRJIT_B200_OP_MAX) { if (vt == DRJIT_B200_VT_INT64) DJB_SC(int64_t, OpMax); else DJB_SC(uint64_t, OpMax); }
            else unsupported();
            break;
        case DRJIT_B200_VT_FLOAT16:
            // min/max need cc >= 90 in the reference (op.cpp:2786-2794); always true here
            if (op == DRJIT_B200_OP_ADD) DJB_SC(__half, OpAdd);
            else if (op == DRJIT_B200_OP_MIN) DJB_SC(__half, OpMin);
            else if (op == DRJIT_B200_OP_MAX) DJB_SC(__half, OpMax);
            else unsupported();
            break;
        case DRJIT_B200_VT_FLOAT32:
            if (op == DRJIT_B200_OP_ADD) DJB_SC(float, OpAdd);
            else if (op == DRJIT_B200_OP_MIN) DJB_SC(float, OpMin);
            else if (op == DRJIT_B200_OP_MAX) DJB_SC(float, OpMax);
            else unsupported();
            break;
        case DRJIT_B200_VT_FLOAT64:
            if (op == DRJIT_B200_OP_ADD) DJB_SC(double, OpAdd);
            else if (op == DRJIT_B200_OP_MIN) DJB_SC(double, OpMin);
            else if (op == DRJIT_B200_OP_MAX) DJB_SC(double, OpMax);
            else unsupported();
            break;
        default:
            unsupported();
    }
#undef DJB_SC
    (void) sign_agnostic;
}

} // namespace djb
