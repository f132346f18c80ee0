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

/// guarded() + the launch hook / kernel history bracket of one primitive call (cuda_ts.cpp:12-47)
template <typename F> int primitive(int kernel_type, uint32_t size, void *stream, F &&f) {
    return guarded([&] {
        djb::CallScope scope(kernel_type, size, S(stream));
        f();
    });
}

} // namespace

extern "C" {

DRJIT_B200_API int drjit_b200_init(void) {
    return guarded([&] { (void) djb::device_props(); });
}

DRJIT_B200_API int drjit_b200_memset_async(void *stream, void *ptr, uint32_t size, uint32_t isize,
                                           const void *src) {
    return primitive(DRJIT_B200_KT_MEMSET, size, stream, [&] { djb::memset_async(S(stream), ptr, size, isize, src); });
}

DRJIT_B200_API int drjit_b200_block_reduce(void *stream, int vt, int op, uint32_t size,
                                           uint32_t block_size, const void *in, void *out) {
    return primitive(DRJIT_B200_KT_BLOCK_REDUCE, size, stream, [&] { djb::block_reduce(S(stream), vt, op, size, block_size, in, out); });
}

DRJIT_B200_API int drjit_b200_block_reduce_bool(void *stream, const uint8_t *values, uint32_t size,
                                                uint8_t *out, int op) {
    return primitive(DRJIT_B200_KT_BLOCK_REDUCE, size, stream, [&] { djb::block_reduce_bool(S(stream), values, size, out, op); });
}

DRJIT_B200_API int drjit_b200_all(void *stream, const uint8_t *values, uint32_t size, int *result) {
    return primitive(DRJIT_B200_KT_BLOCK_REDUCE, size, stream, [&] { *result = djb::all_any(S(stream), values, size, DRJIT_B200_OP_AND); });
}

DRJIT_B200_API int drjit_b200_any(void *stream, const uint8_t *values, uint32_t size, int *result) {
    return primitive(DRJIT_B200_KT_BLOCK_REDUCE, size, stream, [&] { *result = djb::all_any(S(stream), values, size, DRJIT_B200_OP_OR); });
}

DRJIT_B200_API int drjit_b200_reduce_dot(void *stream, int vt, const void *a, const void *b,
                                         uint32_t size, void *out) {
    return primitive(DRJIT_B200_KT_DOT, size, stream, [&] { djb::reduce_dot(S(stream), vt, a, b, size, out); });
}

DRJIT_B200_API int drjit_b200_block_prefix_reduce(void *stream, int vt, int op, uint32_t size,
                                                  uint32_t block_size, int exclusive, int reverse,
                                                  const void *in, void *out) {
    return primitive(DRJIT_B200_KT_BLOCK_PREFIX_REDUCE, size, stream, [&] {
        djb::block_prefix_reduce(S(stream), vt, op, size, block_size, exclusive != 0, reverse != 0,
                                 in, out, nullptr, nullptr);
    });
}

DRJIT_B200_API int drjit_b200_prefix_reduce_carry(void *stream, int vt, int op, uint32_t size,
                                                  int exclusive, int reverse, const void *in,
                                                  void *out, const void *carry_in, void *total_out) {
    return primitive(DRJIT_B200_KT_BLOCK_PREFIX_REDUCE, size, stream, [&] {
        if (size == 0) {
            // empty shard: the running value passes through unchanged (identity without a carry)
            if (total_out && carry_in) {
                DJB_CUDA_CHECK(cudaMemcpyAsync(total_out, carry_in, djb::type_size(vt),
                                               cudaMemcpyDeviceToDevice, S(stream)));
            } else if (total_out) {
                if (djb::type_size(vt) == 0)
                    djb::raise(DRJIT_B200_EUNSUPPORTED, "jit_block_prefix_reduce(): no existing kernel for type=%s, op=%s!",
                               djb::type_name(vt), djb::op_name(op));
                const uint64_t ident = djb::reduce_identity(vt, op);
                djb::memset_async(S(stream), total_out, 1, djb::type_size(vt), &ident);
            }
            return;
        }
        djb::block_prefix_reduce(S(stream), vt, op, size, size, exclusive != 0, reverse != 0, in, out,
                                 carry_in, total_out);
    });
}

DRJIT_B200_API int drjit_b200_compress(void *stream, const uint8_t *in, uint32_t size, uint32_t *out,
                                       uint32_t *count_out) {
    return primitive(DRJIT_B200_KT_COMPRESS, size, stream, [&] {
        const uint32_t count = djb::compress(S(stream), in, size, 0, out, nullptr, true);
        if (count_out) *count_out = count;
    });
}

DRJIT_B200_API int drjit_b200_compress_async(void *stream, const uint8_t *in, uint32_t size,
                                             uint32_t index_base, uint32_t *out, uint32_t *count_dev) {
    return primitive(DRJIT_B200_KT_COMPRESS, size, stream, [&] {
        if (!count_dev)
            djb::raise(DRJIT_B200_EINVAL, "drjit_b200_compress_async(): count_dev must not be NULL!");
        djb::compress(S(stream), in, size, index_base, out, count_dev, false);
    });
}

DRJIT_B200_API int drjit_b200_block_mkperm(void *stream, const uint32_t *values, uint32_t size,
                                           uint32_t block_size, uint32_t bucket_count, uint32_t *perm,
                                           uint32_t *offsets, uint32_t *unique_out) {
    return primitive(DRJIT_B200_KT_MKPERM, size, stream, [&] {
        const uint32_t unique = djb::block_mkperm(S(stream), values, size, block_size, bucket_count, perm, offsets);
        if (unique_out) *unique_out = unique;
    });
}

DRJIT_B200_API int drjit_b200_mkperm_sharded(void *stream, const uint32_t *values, uint32_t size,
                                             uint32_t bucket_count, uint32_t index_base, uint32_t *perm,
                                             uint32_t *hist_dev) {
    return primitive(DRJIT_B200_KT_MKPERM, size, stream, [&] { djb::mkperm_sharded(S(stream), values, size, bucket_count, index_base, perm, hist_dev); });
}

DRJIT_B200_API int drjit_b200_call_reduce(void *stream, const uint32_t *ids, uint32_t size, uint32_t bucket_count,
                                          uint32_t *perm, uint32_t *offsets, uint32_t n_payloads,
                                          const void *const *payload_in, void *const *payload_out,
                                          uint32_t *unique_out) {
    return primitive(DRJIT_B200_KT_MKPERM, size, stream, [&] {
        const uint32_t unique = djb::call_reduce(S(stream), ids, size, bucket_count, perm, offsets, n_payloads, payload_in, payload_out);
        if (unique_out) *unique_out = unique;
    });
}

DRJIT_B200_API int drjit_b200_sort(void *stream, int vt, uint32_t size, int descending, const void *keys,
                                   void *keys_out, uint32_t *index_out) {
    return primitive(DRJIT_B200_KT_SORT, size, stream, [&] { djb::sort(S(stream), vt, size, descending != 0, keys, keys_out, index_out); });
}

DRJIT_B200_API int drjit_b200_poke(void *stream, void *dst, const void *src, uint32_t size) {
    return primitive(DRJIT_B200_KT_POKE, 1, stream, [&] { djb::poke(S(stream), dst, src, size); });
}

DRJIT_B200_API int drjit_b200_aggregate(void *stream, void *dst,
                                        const struct drjit_b200_aggregation_entry *agg, uint32_t size) {
    return primitive(DRJIT_B200_KT_AGGREGATE, size, stream, [&] { djb::aggregate(S(stream), dst, agg, size); });
}

DRJIT_B200_API int drjit_b200_scatter_reduce(void *stream, int vt, int op, int mode, void *target,
                                             uint32_t target_size, const void *value,
                                             const uint32_t *index, const uint8_t *mask, uint32_t size) {
    return primitive(DRJIT_B200_KT_SCATTER_REDUCE, size, stream, [&] {
        djb::scatter_reduce(S(stream), vt, op, mode, target, target_size, value, index, mask, size);
    });
}

DRJIT_B200_API int drjit_b200_scatter_reduce_packet(void *stream, int vt, int op, int mode, void *target,
                                                    uint32_t target_packets, const void *const *values,
                                                    uint32_t count, const uint32_t *index,
                                                    const uint8_t *mask, uint32_t size) {
    return primitive(DRJIT_B200_KT_SCATTER_REDUCE, size, stream, [&] {
        djb::scatter_reduce_packet(S(stream), vt, op, mode, target, target_packets, values, count, index, mask, size);
    });
}

DRJIT_B200_API int drjit_b200_scatter_inc(void *stream, uint32_t *target, uint32_t target_size,
                                          const uint32_t *index, const uint8_t *mask, uint32_t size,
                                          uint32_t *out) {
    return primitive(DRJIT_B200_KT_SCATTER_REDUCE, size, stream, [&] {
        djb::scatter_inc(S(stream), target, target_size, index, mask, size, out);
    });
}

// ---- multi-GPU forms (comm.cu) ---------------------------------------------------------------
#define COMM(c) ((djb::Comm *) (c))

DRJIT_B200_API int drjit_b200_comm_create(uint32_t rank, uint32_t world, size_t bulk_bytes, void **comm_out) {
    return guarded([&] {
        if (!comm_out) djb::raise(DRJIT_B200_EINVAL, "drjit_b200_comm_create(): comm_out is NULL!");
        *comm_out = djb::comm_create(rank, world, bulk_bytes);
    });
}
DRJIT_B200_API int drjit_b200_comm_handle(void *comm, void *handle_out) {
    return guarded([&] {
        if (!comm || !handle_out) djb::raise(DRJIT_B200_EINVAL, "drjit_b200_comm_handle(): NULL argument!");
        djb::comm_handle(COMM(comm), handle_out);
    });
}
DRJIT_B200_API int drjit_b200_comm_connect(void *comm, const void *handles) {
    return guarded([&] {
        if (!comm || !handles) djb::raise(DRJIT_B200_EINVAL, "drjit_b200_comm_connect(): NULL argument!");
        djb::comm_connect_ipc(COMM(comm), handles);
    });
}
DRJIT_B200_API int drjit_b200_comm_connect_local(void **comms, uint32_t world) {
    return guarded([&] {
        if (!comms) djb::raise(DRJIT_B200_EINVAL, "drjit_b200_comm_connect_local(): NULL argument!");
        djb::comm_connect_local((djb::Comm **) comms, world);
    });
}
DRJIT_B200_API int drjit_b200_comm_destroy(void *comm) {
    return guarded([&] { djb::comm_destroy(COMM(comm)); });
}
DRJIT_B200_API int drjit_b200_comm_reduce(void *comm, void *stream, int vt, int op, int fold, uint32_t size,
                                          const void *in, void *out) {
    return primitive(DRJIT_B200_KT_BLOCK_REDUCE, size, stream, [&] {
        djb::comm_reduce(S(stream), COMM(comm), vt, op, (uint32_t) fold, size, in, out);
    });
}
DRJIT_B200_API int drjit_b200_comm_reduce_dot(void *comm, void *stream, int vt, const void *a, const void *b,
                                              uint32_t size, void *out) {
    return primitive(DRJIT_B200_KT_DOT, size, stream, [&] { djb::comm_reduce_dot(S(stream), COMM(comm), vt, a, b, size, out); });
}
DRJIT_B200_API int drjit_b200_comm_all(void *comm, void *stream, const uint8_t *values, uint32_t size, int *result) {
    return primitive(DRJIT_B200_KT_BLOCK_REDUCE, size, stream, [&] {
        *result = djb::comm_all_any(S(stream), COMM(comm), values, size, DRJIT_B200_OP_AND);
    });
}
DRJIT_B200_API int drjit_b200_comm_any(void *comm, void *stream, const uint8_t *values, uint32_t size, int *result) {
    return primitive(DRJIT_B200_KT_BLOCK_REDUCE, size, stream, [&] {
        *result = djb::comm_all_any(S(stream), COMM(comm), values, size, DRJIT_B200_OP_OR);
    });
}
DRJIT_B200_API int drjit_b200_comm_prefix_reduce(void *comm, void *stream, int vt, int op, uint32_t size,
                                                 int exclusive, int reverse, const void *in, void *out,
                                                 void *offset_out, int materialise) {
    return primitive(DRJIT_B200_KT_BLOCK_PREFIX_REDUCE, size, stream, [&] {
        djb::comm_prefix_reduce(S(stream), COMM(comm), vt, op, size, exclusive != 0, reverse != 0, in, out, offset_out,
                                materialise != 0);
    });
}
DRJIT_B200_API int drjit_b200_comm_compress(void *comm, void *stream, const uint8_t *in, uint32_t size,
                                            uint32_t index_base, uint32_t *out, uint32_t *counts_host) {
    return primitive(DRJIT_B200_KT_COMPRESS, size, stream, [&] {
        if (!counts_host) djb::raise(DRJIT_B200_EINVAL, "drjit_b200_comm_compress(): counts_host is NULL!");
        djb::comm_compress(S(stream), COMM(comm), in, size, index_base, out, counts_host);
    });
}
DRJIT_B200_API int drjit_b200_comm_mkperm(void *comm, void *stream, const uint32_t *values, uint32_t size,
                                          uint32_t bucket_count, uint32_t index_base, uint32_t *perm,
                                          uint32_t *hist_dev, uint32_t *rank_base_dev, uint32_t *offsets,
                                          uint32_t *unique_out) {
    return primitive(DRJIT_B200_KT_MKPERM, size, stream, [&] {
        const uint32_t unique = djb::comm_mkperm(S(stream), COMM(comm), values, size, bucket_count, index_base, perm,
                                                 hist_dev, rank_base_dev, offsets);
        if (unique_out) *unique_out = unique;
    });
}
DRJIT_B200_API int drjit_b200_comm_allreduce(void *comm, void *stream, int vt, int op, void *data, uint32_t count) {
    return primitive(DRJIT_B200_KT_PEER_EXCHANGE, count, stream, [&] { djb::comm_allreduce(S(stream), COMM(comm), vt, op, data, count); });
}
DRJIT_B200_API int drjit_b200_comm_allgather(void *comm, void *stream, const void *src, uint32_t bytes, void *dst) {
    return primitive(DRJIT_B200_KT_PEER_EXCHANGE, bytes, stream, [&] { djb::comm_allgather(S(stream), COMM(comm), src, bytes, dst); });
}
DRJIT_B200_API int drjit_b200_comm_fold(void *comm, void *stream, int vt, int op, int fold, const void *src, void *dst) {
    return primitive(DRJIT_B200_KT_PEER_EXCHANGE, 1, stream, [&] {
        djb::comm_fold_scalar(S(stream), COMM(comm), vt, op, (uint32_t) fold, src, dst);
    });
}
#undef COMM

DRJIT_B200_API int drjit_b200_fill_fmix32(void *stream, int kind, void *out, uint64_t start, uint64_t n,
                                          uint32_t xor_, uint32_t and_) {
    return guarded([&] { djb::fill_fmix32(S(stream), kind, out, start, n, xor_, and_); });
}

} // extern "C"
