/*
 * scatter_packet.cu -- the two other scatter forms of the reference's PTX templates as standalone
 * kernels (values / indices already in memory), next to scatter_reduce.cu:
 *
 *  * scatter_reduce_packet: target[index[i] * count + k] op= values[k][i], the packet form that
 *    jitc_cuda_render_scatter_reduce_packet (ext/drjit-core/src/cuda_packet.cpp:168-327) emits for
 *    dr.scatter_reduce / dr.scatter_add of an ArrayNf (film accumulation: RGBA + weight). Where the
 *    hardware has a vector reduction the whole packet (or 16 bytes of it) leaves as ONE instruction:
 *    `red.global.v{2,4}.f32.add`, `red.global.v{2,4,8}.f16.{add,min,max}.noftz` (SASS
 *    REDG.E.ADD.F32x4 ...; cuda_packet.cpp:224-259). Everything else is one RED per component
 *    (:296-325). ReduceMode::Local: one match.any per ELEMENT (the address is shared by the whole
 *    packet, :209-221), peers combine component by component, the lowest lane issues the packet.
 *
 *  * scatter_inc: out[i] = target[index[i]]++ (jitc_cuda_render_scatter_inc,
 *    src/cuda_scatter.cpp:356-393; dr.scatter_inc, the building block of queue compaction). The
 *    reference aggregates per warp (match.any, the leader adds popc(peers), lanes take leader + rank).
 *    Here small counter arrays (<= 2048, i.e. every queue / per-class counter use) are aggregated per
 *    CTA first: a tile of 2048 elements increments counters in shared memory (a coherent warp with
 *    ONE shared-memory atomic), one global atomic per touched counter and tile fetches the base, and
 *    the slots are base + tile-local rank. A single queue counter then sees size / 2048 global atomics
 *    instead of size / 32 on one address (the L2 retires ~1.4 G same-address atomics per second:
 *    DESIGN.md 4.10). Larger counter arrays take the warp-aggregated form. The queue form (no index
 *    array: every active element takes a slot from counter 0) needs no per-element atomic at all: mask
 *    bits, popc, one scan and one global atomic per 8192-element tile (scatter_inc_queue_kernel).
 *    Which lane receives which slot is unspecified in the reference as well; the slots of one counter
 *    are distinct and contiguous from its previous value (tests/test_memop.py:293-316).
 */
#include "common.cuh"
#include "atomic_ops.cuh"
#include "runtime.h"

#include <cuda_fp16.h>

namespace djb {

// ================================================================================================
//  packet scatter-reduce
// ================================================================================================
constexpr uint32_t kPkThreads = 256;
constexpr uint32_t kPkMaxCount = 16;

struct PacketParams {
    void *target;
    const void *values[kPkMaxCount];
    const uint32_t *index;
    const uint8_t *mask;
    uint32_t size, count;
};

/// W consecutive components of a packet -> target; generic: one RED per component
template <typename Op, typename T, uint32_t W> struct PacketRed {
    static __device__ __forceinline__ void apply(T *addr, const T (&v)[W]) {
        #pragma unroll
        for (uint32_t k = 0; k < W; ++k) AtomicOp<Op, T>::apply(addr + k, v[k]);
    }
};
template <> struct PacketRed<OpAdd, float, 2> {
    static __device__ __forceinline__ void apply(float *addr, const float (&v)[2]) {
        asm volatile("red.global.v2.f32.add [%0], {%1, %2};" :: "l"(addr), "f"(v[0]), "f"(v[1]) : "memory");
    }
};
template <> struct PacketRed<OpAdd, float, 4> {
    static __device__ __forceinline__ void apply(float *addr, const float (&v)[4]) {
        asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};"
                     :: "l"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
};
#define DJB_H(i) "h"(__half_as_ushort(v[i]))
#define DJB_PACKET_F16(OP, NAME)                                                                         \
    template <> struct PacketRed<OP, __half, 2> {                                                        \
        static __device__ __forceinline__ void apply(__half *addr, const __half (&v)[2]) {               \
            asm volatile("red.global.v2.f16." NAME ".noftz [%0], {%1, %2};"                              \
                         :: "l"(addr), DJB_H(0), DJB_H(1) : "memory");                                   \
        }                                                                                                \
    };                                                                                                   \
    template <> struct PacketRed<OP, __half, 4> {                                                        \
        static __device__ __forceinline__ void apply(__half *addr, const __half (&v)[4]) {               \
            asm volatile("red.global.v4.f16." NAME ".noftz [%0], {%1, %2, %3, %4};"                      \
                         :: "l"(addr), DJB_H(0), DJB_H(1), DJB_H(2), DJB_H(3) : "memory");               \
        }                                                                                                \
    };                                                                                                   \
    template <> struct PacketRed<OP, __half, 8> {                                                        \
        static __device__ __forceinline__ void apply(__half *addr, const __half (&v)[8]) {               \
            asm volatile("red.global.v8.f16." NAME ".noftz [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"      \
                         :: "l"(addr), DJB_H(0), DJB_H(1), DJB_H(2), DJB_H(3), DJB_H(4), DJB_H(5),       \
                            DJB_H(6), DJB_H(7) : "memory");                                              \
        }                                                                                                \
    };
DJB_PACKET_F16(OpAdd, "add")
DJB_PACKET_F16(OpMin, "min")
DJB_PACKET_F16(OpMax, "max")
#undef DJB_PACKET_F16
#undef DJB_H

/// Lanes of `peers` (equal index, all inside `active`) combine `v`; every lane of `active` calls this
template <typename Op, typename T>
__device__ __forceinline__ T combine_peers(uint32_t active, uint32_t peers, T v) {
    const uint32_t lane = lane_id();
    T acc = v;
    uint32_t rest = peers & ~(1u << lane);
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
    return acc;
}

/// One element (= one packet) per thread and round; the component arrays are read with coalesced
/// scalar loads (32 lanes x sizeof(T) per component), W components leave per reduction.
template <typename T, typename Op, uint32_t W, bool LOCAL>
__global__ void __launch_bounds__(kPkThreads)
scatter_packet_kernel(const PacketParams p) {
    T *target = (T *) p.target;
    const uint64_t gtid = (uint64_t) blockIdx.x * kPkThreads + threadIdx.x,
                   gstride = (uint64_t) gridDim.x * kPkThreads;
    const uint64_t n = p.size;
    // (Local: whole warps stay in the loop, the ballots below name all 32 lanes)
    const uint64_t n_rounded = LOCAL ? (n + gstride - 1) / gstride * gstride : n;
    for (uint64_t i = gtid; i < n_rounded; i += gstride) {
        const bool ok = i < n && (!p.mask || p.mask[i]);
        const uint32_t idx = ok ? p.index[i] : 0u;
        uint32_t active = 0, peers = 0;
        bool leader = true;
        if constexpr (LOCAL) {
            active = __ballot_sync(kFullMask, ok);
            if (ok) {
                peers = __match_any_sync(active, idx);
                leader = (peers & lanemask_lt()) == 0;
            }
        }
        if (!ok)
            continue;
        T *dst = target + (uint64_t) idx * p.count;
        for (uint32_t c = 0; c < p.count; c += W) {
            T v[W];
            #pragma unroll
            for (uint32_t k = 0; k < W; ++k)
                v[k] = ((const T *) p.values[c + k])[i];
            if constexpr (LOCAL) {
                #pragma unroll
                for (uint32_t k = 0; k < W; ++k)
                    v[k] = combine_peers<Op, T>(active, peers, v[k]);
            }
            if (leader)
                PacketRed<Op, T, W>::apply(dst + c, v);
        }
    }
}

/// Fast path of the common case -- Direct mode, 4-byte components, COUNT known at compile time, every
/// array 16-byte aligned: a thread owns four consecutive elements, i.e. one 128-bit load of the indices
/// and one per component (a quarter of the load instructions of the kernel above, no parameter-space
/// indexing), then four packet reductions; the 4 x COUNT register transpose is free.
template <typename T, typename Op, uint32_t COUNT, uint32_t W>
__global__ void __launch_bounds__(kPkThreads)
scatter_packet_vec_kernel(const PacketParams p) {
    static_assert(sizeof(T) == 4 && COUNT % W == 0, "four elements per 128-bit load");
    T *target = (T *) p.target;
    const uint64_t gtid = (uint64_t) blockIdx.x * kPkThreads + threadIdx.x,
                   gstride = (uint64_t) gridDim.x * kPkThreads;
    const uint64_t ngroups = p.size / 4;
    for (uint64_t g = gtid; g < ngroups; g += gstride) {
        const Vec16<uint32_t> iv = ld_stream<uint32_t>(p.index + g * 4);
        Vec16<T> vv[COUNT];
        #pragma unroll
        for (uint32_t k = 0; k < COUNT; ++k)
            vv[k] = ld_stream<T>((const T *) p.values[k] + g * 4);
        const uint32_t m = p.mask ? *reinterpret_cast<const uint32_t *>(p.mask + g * 4) : 0x01010101u;
        #pragma unroll
        for (uint32_t e = 0; e < 4; ++e) {
            if (((m >> (8 * e)) & 0xffu) == 0)
                continue;
            T *dst = target + (uint64_t) iv.v[e] * COUNT;
            #pragma unroll
            for (uint32_t c = 0; c < COUNT; c += W) {
                T v[W];
                #pragma unroll
                for (uint32_t k = 0; k < W; ++k) v[k] = vv[c + k].v[e];
                PacketRed<Op, T, W>::apply(dst + c, v);
            }
        }
    }
    const uint64_t tail = ngroups * 4;
    if (gtid < p.size - tail) {
        const uint64_t i = tail + gtid;
        if (!p.mask || p.mask[i]) {
            T *dst = target + (uint64_t) p.index[i] * COUNT;
            #pragma unroll
            for (uint32_t c = 0; c < COUNT; c += W) {
                T v[W];
                #pragma unroll
                for (uint32_t k = 0; k < W; ++k) v[k] = ((const T *) p.values[c + k])[i];
                PacketRed<Op, T, W>::apply(dst + c, v);
            }
        }
    }
}

template <typename T, typename Op, uint32_t COUNT, uint32_t W>
static void launch_packet_vec(cudaStream_t stream, const PacketParams &p) {
    const DeviceProps &dev = device_props();
    uint32_t grid = (uint32_t) std::min<uint64_t>(ceil_div64(p.size, kPkThreads * 4 * 2), (uint64_t) dev.sm_count * 8 * 4);
    grid = std::max(grid, 1u);
    scatter_packet_vec_kernel<T, Op, COUNT, W><<<grid, kPkThreads, 0, stream>>>(p);
    DJB_POST_LAUNCH();
}

template <typename T, typename Op, uint32_t W, bool LOCAL>
static void launch_packet_w(cudaStream_t stream, const PacketParams &p) {
    const DeviceProps &dev = device_props();
    uint32_t grid = (uint32_t) std::min<uint64_t>(ceil_div64(p.size, kPkThreads * 4), (uint64_t) dev.sm_count * 8 * 4);
    grid = std::max(grid, 1u);
    scatter_packet_kernel<T, Op, W, LOCAL><<<grid, kPkThreads, 0, stream>>>(p);
    DJB_POST_LAUNCH();
}

/// Widest reduction the (type, op) pair has in hardware, limited by the packet size and the
/// alignment of the target (cuda_packet.cpp:229-234: 16 bytes, halved until it divides the count)
template <typename T, typename Op, bool LOCAL>
static void launch_packet(cudaStream_t stream, const PacketParams &p) {
    constexpr bool f32_add = std::is_same<T, float>::value && std::is_same<Op, OpAdd>::value;
    constexpr bool f16 = std::is_same<T, __half>::value;
    if constexpr (f32_add && !LOCAL) {
        bool aligned = ((uintptr_t) p.target % 16) == 0 && ((uintptr_t) p.index % 16) == 0 &&
                       (!p.mask || ((uintptr_t) p.mask % 4) == 0);
        for (uint32_t k = 0; k < p.count; ++k) aligned = aligned && ((uintptr_t) p.values[k] % 16) == 0;
        if (aligned && p.count == 2) return launch_packet_vec<T, Op, 2, 2>(stream, p);
        if (aligned && p.count == 4) return launch_packet_vec<T, Op, 4, 4>(stream, p);
        if (aligned && p.count == 8) return launch_packet_vec<T, Op, 8, 4>(stream, p);
    }
    if constexpr (f32_add || f16) {
        uint32_t w = 16 / sizeof(T);
        while (w > 1 && ((p.count & (w - 1)) != 0 || ((uintptr_t) p.target % (w * sizeof(T))) != 0))
            w /= 2;
        if constexpr (f16) {
            if (w == 8) return launch_packet_w<T, Op, 8, LOCAL>(stream, p);
        }
        if (w == 4) return launch_packet_w<T, Op, 4, LOCAL>(stream, p);
        if (w == 2) return launch_packet_w<T, Op, 2, LOCAL>(stream, p);
    }
    launch_packet_w<T, Op, 1, LOCAL>(stream, p);
}

void scatter_reduce_packet(cudaStream_t stream, int vt, int op, int mode, void *target, uint32_t target_packets,
                           const void *const *values, uint32_t count, const uint32_t *index,
                           const uint8_t *mask, uint32_t size) {
    (void) target_packets;
    if (mode < DRJIT_B200_MODE_AUTO || mode > DRJIT_B200_MODE_PERMUTE)
        raise(DRJIT_B200_EINVAL, "jit_var_scatter_packet(): invalid reduction mode!");
    // cuda_packet.cpp:184-186 (odd packets are not reducible); 16 components is this library's limit
    if (count == 0 || (count & 1u) != 0 || count > kPkMaxCount)
        raise(DRJIT_B200_EINVAL, "jit_var_scatter_packet(): number of elements (%u) not supported by "
              "reduction (must be even and at most %u)", count, kPkMaxCount);
    if (size == 0)
        return;
    if (!target || !values || !index)
        raise(DRJIT_B200_EINVAL, "jit_var_scatter_packet(): null target / values / index array");
    for (uint32_t k = 0; k < count; ++k)
        if (!values[k])
            raise(DRJIT_B200_EINVAL, "jit_var_scatter_packet(): component %u has no data", k);
    PacketParams p{};
    p.target = target; p.index = index; p.mask = mask; p.size = size; p.count = count;
    for (uint32_t k = 0; k < count; ++k) p.values[k] = values[k];

    auto unsupported = [&]() {
        raise(DRJIT_B200_EUNSUPPORTED,
              "jit_var_scatter(): the CUDA backend does not support the requested type of atomic "
              "reduction (%s) for variables of type (%s)", op_name(op), type_name(vt));
    };
    // Local: warp pre-reduction for the types of cuda_packet.cpp:200-207 (no f16); Auto = Direct
    // (at packet granularity the atomics, not the warp reduction, bound the kernel: DESIGN.md 4.10)
    const bool local = mode == DRJIT_B200_MODE_LOCAL;
#define DJB_PK(T, OP) do { if (local) launch_packet<T, OP, true>(stream, p); else launch_packet<T, OP, false>(stream, p); } while (0)
#define DJB_PK_DIRECT(T, OP) launch_packet<T, OP, false>(stream, p)
    switch (vt) {
        case DRJIT_B200_VT_INT32:
        case DRJIT_B200_VT_UINT32:
            if (op == DRJIT_B200_OP_ADD) DJB_PK(uint32_t, OpAdd);
            else if (op == DRJIT_B200_OP_AND) DJB_PK(uint32_t, OpAnd);
            else if (op == DRJIT_B200_OP_OR) DJB_PK(uint32_t, OpOr);
            else if (op == DRJIT_B200_OP_MIN) { if (vt == DRJIT_B200_VT_INT32) DJB_PK(int32_t, OpMin); else DJB_PK(uint32_t, OpMin); }
            else if (op == DRJIT_B200_OP_MAX) { if (vt == DRJIT_B200_VT_INT32) DJB_PK(int32_t, OpMax); else DJB_PK(uint32_t, OpMax); }
            else unsupported();
            break;
        case DRJIT_B200_VT_INT64:
        case DRJIT_B200_VT_UINT64:
            if (op == DRJIT_B200_OP_ADD) DJB_PK(uint64_t, OpAdd);
            else if (op == DRJIT_B200_OP_AND) DJB_PK(uint64_t, OpAnd);
            else if (op == DRJIT_B200_OP_OR) DJB_PK(uint64_t, OpOr);
            else if (op == DRJIT_B200_OP_MIN) { if (vt == DRJIT_B200_VT_INT64) DJB_PK(int64_t, OpMin); else DJB_PK(uint64_t, OpMin); }
            else if (op == DRJIT_B200_OP_MAX) { if (vt == DRJIT_B200_VT_INT64) DJB_PK(int64_t, OpMax); else DJB_PK(uint64_t, OpMax); }
            else unsupported();
            break;
        case DRJIT_B200_VT_FLOAT16:
            if (op == DRJIT_B200_OP_ADD) DJB_PK_DIRECT(__half, OpAdd);
            else if (op == DRJIT_B200_OP_MIN) DJB_PK_DIRECT(__half, OpMin);
            else if (op == DRJIT_B200_OP_MAX) DJB_PK_DIRECT(__half, OpMax);
            else unsupported();
            break;
        case DRJIT_B200_VT_FLOAT32:
            if (op == DRJIT_B200_OP_ADD) DJB_PK(float, OpAdd);
            else if (op == DRJIT_B200_OP_MIN) DJB_PK(float, OpMin);
            else if (op == DRJIT_B200_OP_MAX) DJB_PK(float, OpMax);
            else unsupported();
            break;
        case DRJIT_B200_VT_FLOAT64:
            if (op == DRJIT_B200_OP_ADD) DJB_PK(double, OpAdd);
            else if (op == DRJIT_B200_OP_MIN) DJB_PK(double, OpMin);
            else if (op == DRJIT_B200_OP_MAX) DJB_PK(double, OpMax);
            else unsupported();
            break;
        default:
            unsupported();
    }
#undef DJB_PK
#undef DJB_PK_DIRECT
}

// ================================================================================================
//  scatter_inc
// ================================================================================================
constexpr uint32_t kIncThreads = 256;
constexpr uint32_t kIncPerThread = 8;
constexpr uint32_t kIncTile = kIncThreads * kIncPerThread;
constexpr uint32_t kIncPrivMax = 2048;              // counters aggregated per CTA in shared memory

struct IncParams {
    uint32_t *target;
    const uint32_t *index;      // NULL: every element increments counter 0 (dr.scatter_inc(queue, 0))
    const uint8_t *mask;
    uint32_t *out;
    uint32_t size, target_size;
};

/// Small counter arrays: tile-local ranks from shared-memory atomics, one global atomic per touched
/// counter and tile. Two barriers per tile: `scnt` (counts) is complete after the first; the threads
/// that fetch the bases reset it, so the next tile finds zeros after the second.
__global__ void __launch_bounds__(kIncThreads)
scatter_inc_private_kernel(const IncParams p) {
    __shared__ uint32_t scnt[kIncPrivMax], sbase[kIncPrivMax];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, B = p.target_size;
    for (uint32_t c = tid; c < B; c += kIncThreads) scnt[c] = 0;
    __syncthreads();

    const uint64_t tiles = ((uint64_t) p.size + kIncTile - 1) / kIncTile;
    // indices + activity bits of one tile -> registers
    auto fetch = [&](uint64_t t, uint32_t (&idx)[kIncPerThread], uint32_t &okbits) {
        const uint64_t base = t * kIncTile;
        okbits = 0;
        #pragma unroll
        for (uint32_t j = 0; j < kIncPerThread; ++j) {
            const uint64_t i = base + (uint64_t) j * kIncThreads + tid;
            bool ok = i < p.size && (!p.mask || p.mask[i]);
            idx[j] = ok && p.index ? p.index[i] : 0u;
            ok = ok && idx[j] < B;                  // (out-of-range counters: ignored, slot 0)
            okbits |= (uint32_t) ok << j;
        }
    };
    // The next tile's loads are issued before this tile's barriers and stores: issued after them they
    // queue behind the stores of the other CTAs of the SM and the tile waits for the drain
    // (profiles/r6b_ncu_scatter_packet.md: long_scoreboard).
    uint32_t idx_n[kIncPerThread], okbits_n = 0;
    if (blockIdx.x < tiles) fetch(blockIdx.x, idx_n, okbits_n);
    for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint64_t base = t * kIncTile;
        uint32_t idx[kIncPerThread], loc[kIncPerThread];
        const uint32_t okbits = okbits_n;
        #pragma unroll
        for (uint32_t j = 0; j < kIncPerThread; ++j) idx[j] = idx_n[j];
        if (t + gridDim.x < tiles) fetch(t + gridDim.x, idx_n, okbits_n);
        #pragma unroll
        for (uint32_t j = 0; j < kIncPerThread; ++j) {
            const bool ok = (okbits >> j) & 1u;
            const uint32_t active = __ballot_sync(kFullMask, ok);
            loc[j] = 0;
            if (ok) {
                int same;
                __match_all_sync(active, idx[j], &same);
                if (same) {     // coherent warp (always, for a queue counter): one shared-memory atomic
                    const uint32_t leader = (uint32_t) __ffs(active) - 1;
                    uint32_t old = 0;
                    if (lane == leader) old = atomicAdd(&scnt[idx[j]], (uint32_t) __popc(active));
                    old = __shfl_sync(active, old, leader);
                    loc[j] = old + __popc(active & lanemask_lt());
                } else {
                    loc[j] = atomicAdd(&scnt[idx[j]], 1u);
                }
            }
        }
        __syncthreads();
        for (uint32_t c = tid; c < B; c += kIncThreads) {
            const uint32_t cnt = scnt[c];
            if (cnt) {
                sbase[c] = atomicAdd(p.target + c, cnt);
                scnt[c] = 0;
            }
        }
        __syncthreads();
        #pragma unroll
        for (uint32_t j = 0; j < kIncPerThread; ++j) {
            const uint64_t i = base + (uint64_t) j * kIncThreads + tid;
            if (i < p.size)
                p.out[i] = ((okbits >> j) & 1u) ? sbase[idx[j]] + loc[j] : 0u;
        }
        // (sbase is rewritten only after the first barrier of the next tile, scnt is zero again)
    }
}

/// Queue form (index == NULL: every active element takes a slot from counter 0, dr.scatter_inc(queue, 0,
/// active) -- the stream-compaction use the reference names, jit.h:1137-1138). No per-element atomics at
/// all: a tile of 8192 elements counts its active elements (mask bytes -> bits, popc), one shuffle scan +
/// an 8-entry combine rank them, ONE global atomic per tile fetches the base. A lane owns four consecutive
/// elements per round, so the mask arrives as coalesced 32-bit loads and the slots leave as coalesced
/// 128-bit stores. Two barriers per tile; the per-warp sums and the base are double-buffered so that a
/// fast warp may enter the next tile while a slow one still reads this tile's.
constexpr uint32_t kQueueRounds = 8;
constexpr uint32_t kQueueTile = kIncThreads * 4 * kQueueRounds;

__global__ void __launch_bounds__(kIncThreads)
scatter_inc_queue_kernel(const IncParams p) {
    __shared__ uint32_t wsum[2][kIncThreads / 32], sbase[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool vec = ((uintptr_t) p.out % 16) == 0 && (!p.mask || ((uintptr_t) p.mask % 4) == 0);
    const uint64_t size = p.size, tiles = (size + kQueueTile - 1) / kQueueTile;
    // mask bytes of the lane's 8 x 4 elements of one tile, one 32-bit word per round
    auto fetch = [&](uint64_t t, uint32_t (&m)[kQueueRounds]) {
        const uint64_t base = t * kQueueTile;
        #pragma unroll
        for (uint32_t r = 0; r < kQueueRounds; ++r) {
            const uint64_t e0 = base + (uint64_t) (r * kIncThreads + tid) * 4;
            uint32_t m4 = 0;
            if (vec && e0 + 4 <= size) {
                m4 = p.mask ? *reinterpret_cast<const uint32_t *>(p.mask + e0) : 0x01010101u;
            } else {
                #pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                    if (e0 + k < size && (!p.mask || p.mask[e0 + k]))
                        m4 |= 1u << (8 * k);
            }
            m[r] = m4;
        }
    };
    uint32_t par = 0, m_n[kQueueRounds];
    if (blockIdx.x < tiles) fetch(blockIdx.x, m_n);
    for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x, par ^= 1u) {
        const uint64_t base = t * kQueueTile;
        uint32_t bits = 0;                  // four bits per round: which of the lane's elements are active
        #pragma unroll
        for (uint32_t r = 0; r < kQueueRounds; ++r) {
            const uint32_t m4 = m_n[r];
            const uint32_t nz = (((m4 & 0x7f7f7f7fu) + 0x7f7f7f7fu) | m4) & 0x80808080u;    // bit 7 of each non-zero byte
            const uint32_t b4 = ((nz >> 7) & 1u) | ((nz >> 14) & 2u) | ((nz >> 21) & 4u) | ((nz >> 28) & 8u);
            bits |= b4 << (4 * r);
        }
        // the next tile's mask words travel while this tile is ranked and stored (loads issued after the
        // stores queue behind those of the SM's other CTAs: profiles/r6b_ncu_scatter_packet.md). Without a
        // mask array this is pure arithmetic, but it still has to run: the last tile may be a partial one.
        if (t + gridDim.x < tiles) fetch(t + gridDim.x, m_n);
        const uint32_t count = __popc(bits);
        uint32_t incl = count;
        #pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t o = shfl_up(incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) wsum[par][warp] = incl;
        __syncthreads();
        if (tid == 0) {
            uint32_t total = 0;
            #pragma unroll
            for (uint32_t w = 0; w < kIncThreads / 32; ++w) total += wsum[par][w];
            sbase[par] = total ? atomicAdd(p.target, total) : 0u;
        }
        __syncthreads();
        uint32_t off = sbase[par] + incl - count;
        #pragma unroll
        for (uint32_t w = 0; w < kIncThreads / 32; ++w)
            if (w < warp) off += wsum[par][w];
        #pragma unroll
        for (uint32_t r = 0; r < kQueueRounds; ++r) {
            const uint64_t e0 = base + (uint64_t) (r * kIncThreads + tid) * 4;
            const uint32_t b4 = (bits >> (4 * r)) & 15u;
            uint32_t slot[4];
            #pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                const uint32_t on = (b4 >> k) & 1u;
                slot[k] = on ? off : 0u;
                off += on;
            }
            if (vec && e0 + 4 <= size) {
                Vec16<uint32_t> v;
                #pragma unroll
                for (uint32_t k = 0; k < 4; ++k) v.v[k] = slot[k];
                st_stream<uint32_t>(p.out + e0, v);
            } else {
                #pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                    if (e0 + k < size) p.out[e0 + k] = slot[k];
            }
        }
    }
}

/// Any counter array: the reference's warp-aggregated form (cuda_scatter.cpp:370-388)
__global__ void __launch_bounds__(kIncThreads)
scatter_inc_kernel(const IncParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t gtid = (uint64_t) blockIdx.x * kIncThreads + threadIdx.x,
                   gstride = (uint64_t) gridDim.x * kIncThreads;
    const uint64_t n = p.size, n_rounded = (n + gstride - 1) / gstride * gstride;
    for (uint64_t i = gtid; i < n_rounded; i += gstride) {
        bool ok = i < n && (!p.mask || p.mask[i]);
        const uint32_t idx = ok && p.index ? p.index[i] : 0u;
        ok = ok && idx < p.target_size;
        const uint32_t active = __ballot_sync(kFullMask, ok);
        uint32_t slot = 0;
        if (ok) {
            const uint32_t peers = __match_any_sync(active, idx);
            const uint32_t leader = (uint32_t) __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) old = atomicAdd(p.target + idx, (uint32_t) __popc(peers));
            old = __shfl_sync(active, old, leader);
            slot = old + __popc(peers & lanemask_lt());
        }
        if (i < n)
            p.out[i] = slot;
    }
}

void scatter_inc(cudaStream_t stream, uint32_t *target, uint32_t target_size, const uint32_t *index,
                 const uint8_t *mask, uint32_t size, uint32_t *out) {
    if (size == 0)
        return;
    if (target_size == 0)
        raise(DRJIT_B200_EINVAL, "jit_var_scatter_inc(): the target array is empty");
    if (!target || !out)
        raise(DRJIT_B200_EINVAL, "jit_var_scatter_inc(): null target / output array");
    const DeviceProps &dev = device_props();
    IncParams p{};
    p.target = target; p.index = index; p.mask = mask; p.out = out; p.size = size; p.target_size = target_size;
    if (!index) {
        const uint32_t grid = (uint32_t) std::min<uint64_t>(ceil_div64(size, kQueueTile), (uint64_t) dev.sm_count * 8);
        scatter_inc_queue_kernel<<<grid, kIncThreads, 0, stream>>>(p);
    } else if (target_size <= kIncPrivMax) {
        const uint64_t tiles = ceil_div64(size, kIncTile);
        const uint32_t grid = (uint32_t) std::min<uint64_t>(tiles, (uint64_t) dev.sm_count * 8);
        scatter_inc_private_kernel<<<grid, kIncThreads, 0, stream>>>(p);
    } else {
        const uint32_t grid = (uint32_t) std::min<uint64_t>(ceil_div64(size, kIncThreads * 4), (uint64_t) dev.sm_count * 8 * 4);
        scatter_inc_kernel<<<grid, kIncThreads, 0, stream>>>(p);
    }
    DJB_POST_LAUNCH();
}

} // namespace djb
