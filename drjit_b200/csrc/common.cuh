/*
 * common.cuh -- device-side building blocks shared by all kernels (sm_100a only).
 *
 * Counterpart of the reference's ext/drjit-core/resources/common.h (reduction functors
 * :93-178, status-word helpers :180-255), redesigned: 128-bit streaming loads,
 * redux.sync for 32-bit integer warp reductions, explicit PTX memory-order
 * qualifiers instead of volatile + 64-bit packing tricks.
 */
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ < 1000
#  error "drjit_b200 kernels target sm_100a (B200) only"
#endif

// Timing experiments (phases of a kernel switched off, alternative tile sizes, ...) exist only in
// builds with -DDRJIT_B200_EXPERIMENTS (scripts/ and `make EXPERIMENTS=1`). In the shipped library
// every DJB_DEBUG(...) is the constant 0 and no environment variable is read on the product path.
#if defined(DRJIT_B200_EXPERIMENTS)
#  define DJB_DEBUG(expr) (expr)
#else
#  define DJB_DEBUG(expr) 0u
#endif

namespace djb {

constexpr uint32_t kWarp = 32;
constexpr uint32_t kFullMask = 0xffffffffu;

// ---------------------------------------------------------------------------
//  Type traits: storage type T -> accumulator type A
//  (half accumulates in float: resources/common.h:93-100, src/llvm_red.h:11;
//   u8 is widened to u32 and truncated on store, which is exact modulo 2^8)
// ---------------------------------------------------------------------------
template <typename T> struct Acc { using type = T; };
template <> struct Acc<__half> { using type = float; };
template <> struct Acc<uint8_t> { using type = uint32_t; };
template <typename T> using acc_t = typename Acc<T>::type;

template <typename A, typename T> __device__ __forceinline__ A to_acc(T v) { return (A) v; }
template <> __device__ __forceinline__ float to_acc<float, __half>(__half v) { return __half2float(v); }
template <typename T, typename A> __device__ __forceinline__ T from_acc(A v) { return (T) v; }
template <> __device__ __forceinline__ __half from_acc<__half, float>(float v) { return __float2half_rn(v); }

// ---------------------------------------------------------------------------
//  Reduction functors. identity(): src/var.cpp:2642-2652.
// ---------------------------------------------------------------------------
template <typename A> struct Limits;
template <> struct Limits<uint32_t> { static __device__ __forceinline__ uint32_t lo() { return 0u; } static __device__ __forceinline__ uint32_t hi() { return 0xffffffffu; } };
template <> struct Limits<int32_t>  { static __device__ __forceinline__ int32_t lo() { return (int32_t) 0x80000000; } static __device__ __forceinline__ int32_t hi() { return 0x7fffffff; } };
template <> struct Limits<uint64_t> { static __device__ __forceinline__ uint64_t lo() { return 0ull; } static __device__ __forceinline__ uint64_t hi() { return ~0ull; } };
template <> struct Limits<int64_t>  { static __device__ __forceinline__ int64_t lo() { return (int64_t) 0x8000000000000000ll; } static __device__ __forceinline__ int64_t hi() { return 0x7fffffffffffffffll; } };
template <> struct Limits<float>    { static __device__ __forceinline__ float lo() { return -__int_as_float(0x7f800000); } static __device__ __forceinline__ float hi() { return __int_as_float(0x7f800000); } };
template <> struct Limits<double>   { static __device__ __forceinline__ double lo() { return -__longlong_as_double(0x7ff0000000000000ll); } static __device__ __forceinline__ double hi() { return __longlong_as_double(0x7ff0000000000000ll); } };

__device__ __forceinline__ float  min_(float a, float b)   { return fminf(a, b); }
__device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float  max_(float a, float b)   { return fmaxf(a, b); }
__device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
template <typename A> __device__ __forceinline__ A min_(A a, A b) { return a < b ? a : b; }
template <typename A> __device__ __forceinline__ A max_(A a, A b) { return a < b ? b : a; }

struct OpAdd { static constexpr int id = 1; template <typename A> static __device__ __forceinline__ A identity() { return (A) 0; } template <typename A> static __device__ __forceinline__ A apply(A a, A b) { return a + b; } };
struct OpMul { static constexpr int id = 2; template <typename A> static __device__ __forceinline__ A identity() { return (A) 1; } template <typename A> static __device__ __forceinline__ A apply(A a, A b) { return a * b; } };
struct OpMin { static constexpr int id = 3; template <typename A> static __device__ __forceinline__ A identity() { return Limits<A>::hi(); } template <typename A> static __device__ __forceinline__ A apply(A a, A b) { return min_(a, b); } };
struct OpMax { static constexpr int id = 4; template <typename A> static __device__ __forceinline__ A identity() { return Limits<A>::lo(); } template <typename A> static __device__ __forceinline__ A apply(A a, A b) { return max_(a, b); } };
struct OpAnd { static constexpr int id = 5; template <typename A> static __device__ __forceinline__ A identity() { return (A) ~(A) 0; } template <typename A> static __device__ __forceinline__ A apply(A a, A b) { return a & b; } };
struct OpOr  { static constexpr int id = 6; template <typename A> static __device__ __forceinline__ A identity() { return (A) 0; } template <typename A> static __device__ __forceinline__ A apply(A a, A b) { return a | b; } };

// ---------------------------------------------------------------------------
//  16-byte vectors of T
// ---------------------------------------------------------------------------
template <typename T> struct alignas(16) Vec16 {
    static constexpr uint32_t N = 16 / sizeof(T);
    T v[N];
};

/// Streaming 128-bit load: read-only path, do not allocate in L1 (data is touched once)
template <typename T> __device__ __forceinline__ Vec16<T> ld_stream(const void *ptr) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
    Vec16<T> out;
    *reinterpret_cast<uint4 *>(&out) = r;
    return out;
}

/// Coherent 128-bit load (used when `in` may alias `out`, e.g. in-place scans)
template <typename T> __device__ __forceinline__ Vec16<T> ld_vec(const void *ptr) {
    Vec16<T> out;
    *reinterpret_cast<uint4 *>(&out) = *reinterpret_cast<const uint4 *>(ptr);
    return out;
}

template <typename T> __device__ __forceinline__ void st_stream(void *ptr, const Vec16<T> &v) {
    const uint4 r = *reinterpret_cast<const uint4 *>(&v);
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(ptr), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
}

// ---------------------------------------------------------------------------
//  Memory-order helpers for the decoupled look-back descriptors
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
    uint64_t v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ uint32_t lanemask_le() { uint32_t m; asm("mov.u32 %0, %%lanemask_le;" : "=r"(m)); return m; }

// ---------------------------------------------------------------------------
//  Warp-level reductions. 32-bit integers use redux.sync (one instruction);
//  everything else an xor-butterfly of shuffles.
// ---------------------------------------------------------------------------
template <typename A> __device__ __forceinline__ A shfl_xor(A v, uint32_t m) { return __shfl_xor_sync(kFullMask, v, m); }
template <typename A> __device__ __forceinline__ A shfl_up(A v, uint32_t d) { return __shfl_up_sync(kFullMask, v, d); }
template <typename A> __device__ __forceinline__ A shfl_idx(A v, uint32_t l) { return __shfl_sync(kFullMask, v, l); }

template <typename Op, typename A> struct WarpReduce {
    /// Butterfly over groups of G lanes (G power of two); every lane receives the result
    template <uint32_t G = 32> static __device__ __forceinline__ A run(A v) {
        #pragma unroll
        for (uint32_t m = G / 2; m > 0; m >>= 1)
            v = Op::template apply<A>(v, shfl_xor(v, m));
        return v;
    }
};

#define DJB_REDUX(OP, A, INTRIN)                                                     \
    template <> struct WarpReduce<OP, A> {                                           \
        template <uint32_t G = 32> static __device__ __forceinline__ A run(A v) {    \
            if constexpr (G == 32) {                                                 \
                return (A) INTRIN(kFullMask, v);                                     \
            } else {                                                                 \
                _Pragma("unroll")                                                    \
                for (uint32_t m = G / 2; m > 0; m >>= 1)                             \
                    v = OP::template apply<A>(v, shfl_xor(v, m));                    \
                return v;                                                            \
            }                                                                        \
        }                                                                            \
    };
DJB_REDUX(OpAdd, uint32_t, __reduce_add_sync)
DJB_REDUX(OpAdd, int32_t,  __reduce_add_sync)
DJB_REDUX(OpMin, uint32_t, __reduce_min_sync)
DJB_REDUX(OpMin, int32_t,  __reduce_min_sync)
DJB_REDUX(OpMax, uint32_t, __reduce_max_sync)
DJB_REDUX(OpMax, int32_t,  __reduce_max_sync)
DJB_REDUX(OpAnd, uint32_t, __reduce_and_sync)
DJB_REDUX(OpOr,  uint32_t, __reduce_or_sync)
#undef DJB_REDUX

/// Block-wide reduction (all warps); result valid in thread 0. `smem` holds >= 32 A's.
template <typename Op, typename A, uint32_t Threads>
__device__ __forceinline__ A block_reduce(A v, A *smem, A ident) {
    constexpr uint32_t NW = Threads / 32;
    v = WarpReduce<Op, A>::template run<32>(v);
    if constexpr (NW > 1) {
        const uint32_t w = threadIdx.x >> 5, l = lane_id();
        if (l == 0) smem[w] = v;
        __syncthreads();
        if (w == 0) {
            v = l < NW ? smem[l] : ident;
            v = WarpReduce<Op, A>::template run<32>(v);
        }
    }
    return v;
}

/// fmix32 with the reference test-suite's pre-increment (tests/reductions.cpp:5-13)
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h += 1;
    h ^= h >> 16; h *= 0x85ebca6bu;
    h ^= h >> 13; h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

} // namespace djb
