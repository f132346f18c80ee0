/*
 * drjit_b200_thread_state.h -- C++ adapter with the exact method signatures of the reference's
 * CUDAThreadState (ext/drjit-core/src/cuda_ts.h:10-50, virtuals declared in
 * src/internal.h:902-962), implemented on top of the C ABI in drjit_b200.h.
 *
 * A Dr.Jit-Core maintainer can either (a) derive CUDAThreadState's primitive methods from this
 * class, or (b) paste the six one-line bodies below into src/cuda_ts.cpp in place of the
 * existing ones. Error behaviour mirrors the reference:
 *   jitc_raise()  -> std::runtime_error   (DRJIT_B200_EINVAL / EUNSUPPORTED)
 *   jitc_fail()   -> message on stderr + abort()   (DRJIT_B200_ECUDA / EFATAL)
 *
 * The enum classes re-declare the numeric values of include/drjit-core/jit.h so that this
 * header also compiles stand-alone; inside drjit-core define DRJIT_B200_USE_JIT_H before
 * including it to use the real `VarType` / `ReduceOp` / `AggregationEntry`.
 */
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "drjit_b200.h"

#if !defined(DRJIT_B200_USE_JIT_H)
enum class VarType : uint32_t {   // jit.h:597-611
    Void, Bool, BaseInt, Int8, UInt8, Int16, UInt16, Int32, UInt32, Int64, UInt64,
    Pointer, BaseFloat, Float16, Float32, Float64, Count
};
enum class ReduceOp : uint32_t {  // jit.h:990-1014
    Identity, Add, Mul, Min, Max, And, Or, Count
};
enum class ReduceMode : uint32_t { // jit.h:1017-1066
    Auto, Direct, Local, NoConflicts, Expand, Permute
};
struct AggregationEntry {          // jit.h:2435-2443
    int16_t size;
    uint16_t resource_kind;
    uint32_t offset;
    const void *src;
};
#endif

namespace drjit_b200 {

/// Translate a status code the way the reference reports errors
inline void check(int status) {
    if (status == DRJIT_B200_OK)
        return;
    if (status == DRJIT_B200_EINVAL || status == DRJIT_B200_EUNSUPPORTED)
        throw std::runtime_error(drjit_b200_last_error());   // jitc_raise(), src/log.cpp:166-169
    fprintf(stderr, "Critical Dr.Jit compiler failure: %s\n", drjit_b200_last_error());
    abort();                                                  // jitc_fail(), src/log.cpp:195-198
}

/// Drop-in for the primitive part of `struct CUDAThreadState : ThreadState`
struct ThreadState {
    /// The stream all work is enqueued on (ThreadState::stream, src/internal.h:890)
    void *stream = nullptr;

    explicit ThreadState(void *stream_ = nullptr) : stream(stream_) { }

    /// src/cuda_ts.cpp:129
    void memset_async(void *ptr, uint32_t size, uint32_t isize, const void *src) {
        check(drjit_b200_memset_async(stream, ptr, size, isize, src));
    }

    /// src/cuda_ts.cpp:195
    void block_reduce(VarType vt, ReduceOp op, uint32_t size, uint32_t block_size,
                      const void *in, void *out) {
        check(drjit_b200_block_reduce(stream, (int) vt, (int) op, size, block_size, in, out));
    }

    /// src/init.cpp:919 (the caller's buffer is NOT padded, unlike the reference)
    void block_reduce_bool(uint8_t *values, uint32_t size, uint8_t *out, ReduceOp op) {
        check(drjit_b200_block_reduce_bool(stream, values, size, out, (int) op));
    }

    /// src/cuda_ts.cpp:530
    void block_prefix_reduce(VarType vt, ReduceOp op, uint32_t size, uint32_t block_size,
                             bool exclusive, bool reverse, const void *in, void *out) {
        check(drjit_b200_block_prefix_reduce(stream, (int) vt, (int) op, size, block_size,
                                             exclusive, reverse, in, out));
    }

    /// src/cuda_ts.cpp:354
    void reduce_dot(VarType vt, const void *ptr_1, const void *ptr_2, uint32_t size, void *out) {
        check(drjit_b200_reduce_dot(stream, (int) vt, ptr_1, ptr_2, size, out));
    }

    /// src/cuda_ts.cpp:683 (synchronous)
    uint32_t compress(const uint8_t *in, uint32_t size, uint32_t *out) {
        uint32_t count = 0;
        check(drjit_b200_compress(stream, in, size, out, &count));
        return count;
    }

    /// src/cuda_ts.cpp:788 (waits for the bucket table only)
    uint32_t block_mkperm(const uint32_t *values, uint32_t size, uint32_t block_size,
                          uint32_t bucket_count, uint32_t *perm, uint32_t *offsets) {
        uint32_t unique = 0;
        check(drjit_b200_block_mkperm(stream, values, size, block_size, bucket_count, perm,
                                      offsets, &unique));
        return unique;
    }

    /// src/cuda_ts.cpp:988
    void poke(void *dst, const void *src, uint32_t size) {
        check(drjit_b200_poke(stream, dst, src, size));
    }

    /// src/cuda_ts.cpp:1008
    void aggregate(void *dst, AggregationEntry *agg, uint32_t size) {
        static_assert(sizeof(AggregationEntry) == sizeof(drjit_b200_aggregation_entry),
                      "AggregationEntry layout mismatch");
        check(drjit_b200_aggregate(stream, dst, (const drjit_b200_aggregation_entry *) agg, size));
    }

    /// jitc_all / jitc_any, src/util.cpp:177-211
    bool all(uint8_t *values, uint32_t size) {
        int r = 0;
        check(drjit_b200_all(stream, values, size, &r));
        return r != 0;
    }
    bool any(uint8_t *values, uint32_t size) {
        int r = 0;
        check(drjit_b200_any(stream, values, size, &r));
        return r != 0;
    }

    /// Standalone form of the scatter-reduce the reference JIT-emits (src/cuda_scatter.cpp:246-354)
    void scatter_reduce(VarType vt, ReduceOp op, ReduceMode mode, void *target, uint32_t target_size,
                        const void *value, const uint32_t *index, const uint8_t *mask, uint32_t size) {
        check(drjit_b200_scatter_reduce(stream, (int) vt, (int) op, (int) mode, target, target_size,
                                        value, index, mask, size));
    }

    /// Standalone form of the packet scatter-reduce template (src/cuda_packet.cpp:168-327):
    /// target[index[i] * count + k] op= values[k][i]
    void scatter_reduce_packet(VarType vt, ReduceOp op, ReduceMode mode, void *target, uint32_t target_packets,
                               const void *const *values, uint32_t count, const uint32_t *index,
                               const uint8_t *mask, uint32_t size) {
        check(drjit_b200_scatter_reduce_packet(stream, (int) vt, (int) op, (int) mode, target, target_packets,
                                               values, count, index, mask, size));
    }

    /// Standalone form of the scatter_inc template (src/cuda_scatter.cpp:356-393): out[i] = target[index[i]]++;
    /// index == nullptr: every active element takes a slot from target[0]
    void scatter_inc(uint32_t *target, uint32_t target_size, const uint32_t *index, const uint8_t *mask,
                     uint32_t size, uint32_t *out) {
        check(drjit_b200_scatter_inc(stream, target, target_size, index, mask, size, out));
    }
};

} // namespace drjit_b200
