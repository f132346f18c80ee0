/*
 * atomic_ops.cuh -- per-element reductions into global memory without a return value (RED.E.<op>),
 * shared by the scatter kernels (scatter_reduce.cu, scatter_packet.cu). Follows the reference's
 * PTX templates: src/cuda_scatter.cpp:74-106 (float min/max through integer atomics), :291-332
 * (two-wide f16 forms with an identity partner).
 */
#pragma once

#include "common.cuh"

namespace djb {

// ---- per-element atomic, no return value (compiles to RED.E.<op>) -----------------------
template <typename Op, typename T> struct AtomicOp;

#define DJB_ATOMIC(OP, T, EXPR)                                                            \
    template <> struct AtomicOp<OP, T> {                                                   \
        static __device__ __forceinline__ void apply(T *addr, T v) { EXPR; }               \
    };
DJB_ATOMIC(OpAdd, uint32_t, atomicAdd(addr, v))
DJB_ATOMIC(OpAdd, uint64_t, atomicAdd((unsigned long long *) addr, (unsigned long long) v))
DJB_ATOMIC(OpAdd, float, atomicAdd(addr, v))
DJB_ATOMIC(OpAdd, double, atomicAdd(addr, v))
DJB_ATOMIC(OpMin, uint32_t, atomicMin(addr, v))
DJB_ATOMIC(OpMin, int32_t, atomicMin(addr, v))
DJB_ATOMIC(OpMin, uint64_t, atomicMin((unsigned long long *) addr, (unsigned long long) v))
DJB_ATOMIC(OpMin, int64_t, atomicMin((long long *) addr, (long long) v))
DJB_ATOMIC(OpMax, uint32_t, atomicMax(addr, v))
DJB_ATOMIC(OpMax, int32_t, atomicMax(addr, v))
DJB_ATOMIC(OpMax, uint64_t, atomicMax((unsigned long long *) addr, (unsigned long long) v))
DJB_ATOMIC(OpMax, int64_t, atomicMax((long long *) addr, (long long) v))
DJB_ATOMIC(OpAnd, uint32_t, atomicAnd(addr, v))
DJB_ATOMIC(OpAnd, uint64_t, atomicAnd((unsigned long long *) addr, (unsigned long long) v))
DJB_ATOMIC(OpOr, uint32_t, atomicOr(addr, v))
DJB_ATOMIC(OpOr, uint64_t, atomicOr((unsigned long long *) addr, (unsigned long long) v))
// float min/max: values with a clear sign bit order like signed ints, values with the sign bit set
// like reversed unsigned ints. The branch tests the BIT PATTERN like the reference's
// `setp.ge.s32` (cuda_scatter.cpp:98-105), not the float value: -0.0 (0x80000000) must take the
// unsigned path -- as a signed int it is INT_MIN, which atomicMin would store over any more
// negative target and atomicMax would never store.
DJB_ATOMIC(OpMin, float, if (__float_as_int(v) >= 0) atomicMin((int *) addr, __float_as_int(v));
                         else atomicMax((unsigned *) addr, __float_as_uint(v)))
DJB_ATOMIC(OpMax, float, if (__float_as_int(v) >= 0) atomicMax((int *) addr, __float_as_int(v));
                         else atomicMin((unsigned *) addr, __float_as_uint(v)))
DJB_ATOMIC(OpMin, double, if (__double_as_longlong(v) >= 0) atomicMin((long long *) addr, __double_as_longlong(v));
                          else atomicMax((unsigned long long *) addr, (unsigned long long) __double_as_longlong(v)))
DJB_ATOMIC(OpMax, double, if (__double_as_longlong(v) >= 0) atomicMax((long long *) addr, __double_as_longlong(v));
                          else atomicMin((unsigned long long *) addr, (unsigned long long) __double_as_longlong(v)))
#undef DJB_ATOMIC

// f16 min/max: the hardware has only the two-wide form (sm_90+, `red.global.v2.f16.{min,max}`);
// the other half of the aligned pair receives the identity (+inf / -inf), exactly as
// cuda_scatter.cpp:307-332 does it. Like there, the partner of the last element of an
// odd-sized target lies 2 bytes past its end (inside the same 4-byte word).
template <bool IS_MIN>
__device__ __forceinline__ void red_f16_minmax(__half *addr, __half v) {
    const uint16_t ident = IS_MIN ? 0x7c00u : 0xfc00u, bits = __half_as_ushort(v);
    const bool even = (((uintptr_t) addr) & 2u) == 0;
    const uint16_t lo = even ? bits : ident, hi = even ? ident : bits;
    const uintptr_t base = ((uintptr_t) addr) & ~(uintptr_t) 2;
    if (IS_MIN)
        asm volatile("red.global.v2.f16.min.noftz [%0], {%1, %2};" :: "l"(base), "h"(lo), "h"(hi) : "memory");
    else
        asm volatile("red.global.v2.f16.max.noftz [%0], {%1, %2};" :: "l"(base), "h"(lo), "h"(hi) : "memory");
}
// f16 add: atomicAdd(__half *) compiles to a compare-and-swap loop; the packed form with a zero
// partner is a native fire-and-forget reduction (cuda_scatter.cpp:291-306 makes the same choice)
template <> struct AtomicOp<OpAdd, __half> {
    static __device__ __forceinline__ void apply(__half *addr, __half v) {
        const uint32_t bits = __half_as_ushort(v);
        const uint32_t packed = (((uintptr_t) addr) & 2u) ? bits << 16 : bits;
        asm volatile("red.global.add.noftz.f16x2 [%0], %1;"
                     :: "l"(((uintptr_t) addr) & ~(uintptr_t) 2), "r"(packed) : "memory");
    }
};
template <> struct AtomicOp<OpMin, __half> {
    static __device__ __forceinline__ void apply(__half *addr, __half v) { red_f16_minmax<true>(addr, v); }
};
template <> struct AtomicOp<OpMax, __half> {
    static __device__ __forceinline__ void apply(__half *addr, __half v) { red_f16_minmax<false>(addr, v); }
};

} // namespace djb
