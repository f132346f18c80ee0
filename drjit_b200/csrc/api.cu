/*
 * api.cu -- the extern "C" boundary (include/drjit_b200.h): argument marshalling and the
 * translation of djb::Error into status codes + a thread-local message. Counterpart of the
 * "lock + forward" wrappers in ext/drjit-core/src/api.cpp:1303-1350.
 */
#include "runtime.h"

#include <cstdio>
#include <cstring>

extern thread_local char t_last_error[1024];

namespace {

template <typename F> int guarded(F &&f) {
    try {
        f();
        t_last_error[0] = '\0';
        return DRJIT_B200_OK;
    } catch (const djb::Error &e) {
        snprintf(t_last_error, sizeof(t_last_error), "%s", e.what());
        return e.code;
    } catch (const std::exception &e) {
        snprintf(t_last_error, sizeof(t_last_error), "%s", e.what());
        return DRJIT_B200_EFATAL;
    }
}

inline cudaStream_t S(void *stream) { return (cudaStream_t) stream; }

} // namespace

extern "C" {

DRJIT_B200_API int drjit_b200_init(void) {
    return guarded([&] { (void) djb::device_props(); });
}

DRJIT_B200_API int drjit_b200_memset_async(void *stream, void *ptr, uint32_t size, uint32_t isize,
                                           const void *src) {
    return guarded([&] { djb::memset_async(S(stream), ptr, size, isize, src); });
}

DRJIT_B200_API int drjit_b200_block_reduce(void *stream, int vt, int op, uint32_t size,
                                           uint32_t block_size, const void *in, void *out) {
    return guarded([&] { djb::block_reduce(S(stream), vt, op, size, block_size, in, out); });
}

DRJIT_B200_API int drjit_b200_block_reduce_bool(void *stream, const uint8_t *values, uint32_t size,
                                                uint8_t *out, int op) {
    return guarded([&] { djb::block_reduce_bool(S(stream), values, size, out, op); });
}

DRJIT_B200_API int drjit_b200_all(void *stream, const uint8_t *values, uint32_t size, int *result) {
    return guarded([&] { *result = djb::all_any(S(stream), values, size, DRJIT_B200_OP_AND); });
}

DRJIT_B200_API int drjit_b200_any(void *stream, const uint8_t *values, uint32_t size, int *result) {
    return guarded([&] { *result = djb::all_any(S(stream), values, size, DRJIT_B200_OP_OR); });
}

DRJIT_B200_API int drjit_b200_reduce_dot(void *stream, int vt, const void *a, const void *b,
                                         uint32_t size, void *out) {
    return guarded([&] { djb::reduce_dot(S(stream), vt, a, b, size, out); });
}

DRJIT_B200_API int drjit_b200_block_prefix_reduce(void *stream, int vt, int op, uint32_t size,
                                                  uint32_t block_size, int exclusive, int reverse,
                                                  const void *in, void *out) {
    return guarded([&] {
        djb::block_prefix_reduce(S(stream), vt, op, size, block_size, exclusive != 0, reverse != 0,
                                 in, out, nullptr, nullptr);
    });
}

DRJIT_B200_API int drjit_b200_prefix_reduce_carry(void *stream, int vt, int op, uint32_t size,
                                                  int exclusive, int reverse, const void *in,
                                                  void *out, const void *carry_in, void *total_out) {
    return guarded([&] {
        if (size == 0) {
            // empty shard: the running value passes through unchanged
            if (total_out && carry_in)
                DJB_CUDA_CHECK(cudaMemcpyAsync(total_out, carry_in, djb::type_size(vt),
                                               cudaMemcpyDeviceToDevice, S(stream)));
            return;
        }
        djb::block_prefix_reduce(S(stream), vt, op, size, size, exclusive != 0, reverse != 0, in, out,
                                 carry_in, total_out);
    });
}

DRJIT_B200_API int drjit_b200_compress(void *stream, const uint8_t *in, uint32_t size, uint32_t *out,
                                       uint32_t *count_out) {
    return guarded([&] {
        const uint32_t count = djb::compress(S(stream), in, size, 0, out, nullptr, true);
        if (count_out) *count_out = count;
    });
}

DRJIT_B200_API int drjit_b200_compress_async(void *stream, const uint8_t *in, uint32_t size,
                                             uint32_t index_base, uint32_t *out, uint32_t *count_dev) {
    return guarded([&] {
        if (!count_dev)
            djb::raise(DRJIT_B200_EINVAL, "drjit_b200_compress_async(): count_dev must not be NULL!");
        djb::compress(S(stream), in, size, index_base, out, count_dev, false);
    });
}

DRJIT_B200_API int drjit_b200_block_mkperm(void *stream, const uint32_t *values, uint32_t size,
                                           uint32_t block_size, uint32_t bucket_count, uint32_t *perm,
                                           uint32_t *offsets, uint32_t *unique_out) {
    return guarded([&] {
        const uint32_t unique = djb::block_mkperm(S(stream), values, size, block_size, bucket_count, perm, offsets);
        if (unique_out) *unique_out = unique;
    });
}

DRJIT_B200_API int drjit_b200_mkperm_sharded(void *stream, const uint32_t *values, uint32_t size,
                                             uint32_t bucket_count, uint32_t index_base, uint32_t *perm,
                                             uint32_t *hist_dev) {
    return guarded([&] { djb::mkperm_sharded(S(stream), values, size, bucket_count, index_base, perm, hist_dev); });
}

DRJIT_B200_API int drjit_b200_poke(void *stream, void *dst, const void *src, uint32_t size) {
    return guarded([&] { djb::poke(S(stream), dst, src, size); });
}

DRJIT_B200_API int drjit_b200_aggregate(void *stream, void *dst,
                                        const struct drjit_b200_aggregation_entry *agg, uint32_t size) {
    return guarded([&] { djb::aggregate(S(stream), dst, agg, size); });
}

DRJIT_B200_API int drjit_b200_scatter_reduce(void *stream, int vt, int op, int mode, void *target,
                                             uint32_t target_size, const void *value,
                                             const uint32_t *index, const uint8_t *mask, uint32_t size) {
    return guarded([&] {
        djb::scatter_reduce(S(stream), vt, op, mode, target, target_size, value, index, mask, size);
    });
}

DRJIT_B200_API int drjit_b200_fill_fmix32(void *stream, int kind, void *out, uint64_t start, uint64_t n,
                                          uint32_t xor_, uint32_t and_) {
    return guarded([&] { djb::fill_fmix32(S(stream), kind, out, start, n, xor_, and_); });
}

} // extern "C"
