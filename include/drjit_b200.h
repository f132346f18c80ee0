/*
 * drjit_b200.h -- C ABI of the B200-native (sm_100a) data-parallel primitive layer
 * that replaces the CUDA primitives behind Dr.Jit-Core's ThreadState seam.
 *
 * Every entry point below is the plain-C form of one reference interface; the
 * citation (relative to the reference tree, ext/drjit-core/) names the function
 * it replaces. Compared to the reference signatures each call takes one extra
 * leading argument: the CUDA stream (`CUstream`/`cudaStream_t` passed as void*)
 * that the reference keeps in `ThreadState::stream` (src/internal.h:882-900).
 *
 * Conventions
 *  - All data pointers are device pointers owned by the caller, except where a
 *    parameter is documented as "host" (pinned or pageable host memory).
 *  - `vt` is a Dr.Jit `VarType` value, `op` a `ReduceOp` value
 *    (include/drjit-core/jit.h:597-611 and :990-1014); the numeric values are
 *    re-declared below so that this header is self-contained.
 *  - Return value: 0 on success, a negative DRJIT_B200_E* code otherwise. After an
 *    error, drjit_b200_last_error() returns a thread-local message with the same
 *    wording as the reference's jitc_raise()/jitc_fail() text where one exists.
 *    The C++ adapter (drjit_b200_thread_state.h) turns these codes back into
 *    std::runtime_error / abort() exactly like the reference does.
 *  - Everything is asynchronous on `stream` unless stated otherwise.
 *  - There is no CPU fallback: every function fails with DRJIT_B200_ECUDA when no
 *    sm_100 device/context is usable.
 */
#ifndef DRJIT_B200_H
#define DRJIT_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__cplusplus)
extern "C" {
#endif

#if defined(DRJIT_B200_BUILD)
#  define DRJIT_B200_API __attribute__((visibility("default")))
#else
#  define DRJIT_B200_API
#endif

/* ---- enums (values identical to include/drjit-core/jit.h) ---------------- */
enum drjit_b200_var_type {          /* jit.h:597-611 */
    DRJIT_B200_VT_VOID = 0, DRJIT_B200_VT_BOOL = 1, DRJIT_B200_VT_INT8 = 3,
    DRJIT_B200_VT_UINT8 = 4, DRJIT_B200_VT_INT16 = 5, DRJIT_B200_VT_UINT16 = 6,
    DRJIT_B200_VT_INT32 = 7, DRJIT_B200_VT_UINT32 = 8, DRJIT_B200_VT_INT64 = 9,
    DRJIT_B200_VT_UINT64 = 10, DRJIT_B200_VT_FLOAT16 = 13,
    DRJIT_B200_VT_FLOAT32 = 14, DRJIT_B200_VT_FLOAT64 = 15
};

enum drjit_b200_reduce_op {         /* jit.h:990-1014 */
    DRJIT_B200_OP_IDENTITY = 0, DRJIT_B200_OP_ADD = 1, DRJIT_B200_OP_MUL = 2,
    DRJIT_B200_OP_MIN = 3, DRJIT_B200_OP_MAX = 4, DRJIT_B200_OP_AND = 5,
    DRJIT_B200_OP_OR = 6
};

enum drjit_b200_reduce_mode {       /* jit.h:1017-1066 */
    DRJIT_B200_MODE_AUTO = 0, DRJIT_B200_MODE_DIRECT = 1, DRJIT_B200_MODE_LOCAL = 2,
    DRJIT_B200_MODE_NO_CONFLICTS = 3, DRJIT_B200_MODE_EXPAND = 4,
    DRJIT_B200_MODE_PERMUTE = 5
};

enum drjit_b200_status {
    DRJIT_B200_OK = 0,
    DRJIT_B200_EINVAL = -1,        /* reference: jitc_raise() -> std::runtime_error */
    DRJIT_B200_EUNSUPPORTED = -2,  /* reference: jitc_raise("no existing kernel for type=..") */
    DRJIT_B200_ECUDA = -3,         /* reference: cuda_check() -> jitc_fail() -> abort() */
    DRJIT_B200_EFATAL = -4         /* reference: jitc_fail() -> abort() (e.g. bucket_count == 0) */
};

/* AggregationEntry, include/drjit-core/jit.h:2435-2443 (16 bytes, 16-byte aligned) */
struct drjit_b200_aggregation_entry {
    int16_t size;            /* >0: literal of that many bytes held in `src`; <0: copy |size| bytes from `src` */
    uint16_t resource_kind;  /* ignored by the kernel */
    uint32_t offset;         /* byte offset into the destination block */
    const void *src;
};

/* ---- library / device management ---------------------------------------- */

/* Thread-local text of the last error raised on the calling thread ("" if none). */
DRJIT_B200_API const char *drjit_b200_last_error(void);

/* Library version string, e.g. "drjit-b200 0.1 (sm_100a)". */
DRJIT_B200_API const char *drjit_b200_version(void);

/* Eagerly create the per-device state (SM count, opt-in shared memory, scratch arena)
 * for the device that is current on the calling thread; otherwise created lazily by
 * the first primitive. Replaces the per-device part of jitc_cuda_init(),
 * src/cuda_core.cpp:266-539 -- there is no PTX to JIT: kernels are sm_100a SASS. */
DRJIT_B200_API int drjit_b200_init(void);

/* Release scratch arenas and pinned staging buffers of all devices
 * (counterpart of jitc_cuda_shutdown(), src/cuda_core.cpp:541-582). */
DRJIT_B200_API int drjit_b200_shutdown(void);

/* Scratch memory hooks. The reference obtains temporaries from its stream-ordered
 * caching allocator, jitc_malloc()/jitc_free() (src/malloc.cpp:102-200). When linked
 * into drjit-core, pass thin wrappers around those two functions; `pinned != 0`
 * asks for host-pinned ("shared", malloc.cpp:182-183) memory. With no hooks
 * installed the library keeps a grow-only arena per (device, stream). */
typedef void *(*drjit_b200_malloc_fn)(size_t size, int pinned, void *user);
typedef void (*drjit_b200_free_fn)(void *ptr, void *user);
DRJIT_B200_API int drjit_b200_set_allocator(drjit_b200_malloc_fn malloc_fn,
                                            drjit_b200_free_fn free_fn, void *user);

/* Number of kernels launched by this library on the calling thread since the last
 * call with reset != 0 (KernelHistory-style accounting, src/cuda_ts.cpp:19-46). */
DRJIT_B200_API uint64_t drjit_b200_launch_count(int reset);

/* ---- the ThreadState seam (src/internal.h:902-962, src/cuda_ts.h:10-50) -- */

/* CUDAThreadState::memset_async, src/cuda_ts.cpp:129-183 (+ fill_64, resources/misc.cuh:27-31).
 * Fills `size` elements of `isize` in {1,2,4,8} bytes with the pattern at host pointer `src`. */
DRJIT_B200_API int drjit_b200_memset_async(void *stream, void *ptr, uint32_t size,
                                           uint32_t isize, const void *src);

/* CUDAThreadState::block_reduce, src/cuda_ts.cpp:195-352 (kernel resources/block_reduce.cuh:93-219).
 * out[b] = op-reduction of in[b*block_size .. min((b+1)*block_size, size)).
 * `block_size == size` is the full reduction behind jit_reduce / dr.sum (src/util.cpp:42-45).
 * size == 0: no-op. block_size == 0 or > size: EINVAL. block_size == 1: device copy.
 * Types: u8,i32,u32,i64,u64,f16,f32,f64; And/Or only for integer types. */
DRJIT_B200_API int drjit_b200_block_reduce(void *stream, int vt, int op, uint32_t size,
                                           uint32_t block_size, const void *in, void *out);

/* ThreadState::block_reduce_bool, src/init.cpp:919-939 (dr.all / dr.any, src/util.cpp:153-175).
 * Writes the 4-byte partial (four packed bools) to `out`; `op` is And or Or. Unlike the
 * reference this does NOT write padding bytes past values[size). */
DRJIT_B200_API int drjit_b200_block_reduce_bool(void *stream, const uint8_t *values,
                                                uint32_t size, uint8_t *out, int op);

/* jitc_all / jitc_any, src/util.cpp:177-211: synchronous, result (0/1) in *result (host). */
DRJIT_B200_API int drjit_b200_all(void *stream, const uint8_t *values, uint32_t size, int *result);
DRJIT_B200_API int drjit_b200_any(void *stream, const uint8_t *values, uint32_t size, int *result);

/* CUDAThreadState::reduce_dot, src/cuda_ts.cpp:354-398 (kernel resources/reduce_2.cuh:12-76).
 * out[0] = sum_i a[i]*b[i] with FMA; vt in {f16,f32,f64}. */
DRJIT_B200_API int drjit_b200_reduce_dot(void *stream, int vt, const void *a, const void *b,
                                         uint32_t size, void *out);

/* CUDAThreadState::block_prefix_reduce, src/cuda_ts.cpp:530-681
 * (kernel resources/block_prefix_reduce.cuh:46-214). Argument order is the EFFECTIVE
 * positional order of jit_block_prefix_reduce: (size, block_size) -- see
 * src/api.cpp:1331-1337 -> src/util.cpp:55-61. In-place (out == in) is allowed. */
DRJIT_B200_API int drjit_b200_block_prefix_reduce(void *stream, int vt, int op, uint32_t size,
                                                  uint32_t block_size, int exclusive,
                                                  int reverse, const void *in, void *out);

/* CUDAThreadState::compress, src/cuda_ts.cpp:683-763 (kernels resources/compress.cuh:23-156).
 * Writes the ascending indices of non-zero mask bytes to out[0..count) and returns the
 * count through *count_out (host). SYNCHRONOUS like the reference (cuda_ts.cpp:759).
 * Unlike the reference it never writes to `in` (no zero padding of the mask tail). */
DRJIT_B200_API int drjit_b200_compress(void *stream, const uint8_t *in, uint32_t size,
                                       uint32_t *out, uint32_t *count_out);

/* CUDAThreadState::block_mkperm, src/cuda_ts.cpp:788-975 (kernels resources/mkperm.cuh:14-499).
 * perm (device, `size` entries): permutation sorting each group of `block_size` keys by key.
 * offsets (host-pinned, 4*bucket_count+1 entries, may be NULL): when block_size == size,
 * quadruples {bucket id, start, size, 0} of the non-empty buckets in ascending id order and
 * the unique count at offsets[4*bucket_count]; the count is also returned in *unique_out.
 * As in the reference (cuda_ts.cpp:953-967) the call waits only until the bucket table is
 * valid; `perm` is complete in stream order. bucket_count == 0: EFATAL.
 * Stability (include/drjit-core/jit.h:2404-2406): the permutation is stable -- equal keys keep their
 * input order, bit-identical to the LLVM backend -- whenever the reference's "tiny" variant would
 * apply (bucket_count * 4 bytes * 32 warps fit into shared memory: <= 1816 buckets on a B200), at
 * every input size; dr.sort / dr.argsort (LSD radix passes with 256 buckets) depend on it. Inputs
 * below 2^18 keys and block_size < size are stable up to 7264 buckets. Beyond that the keys of a
 * bucket appear in unspecified order, like in the reference's "small" / "large" variants. */
DRJIT_B200_API int drjit_b200_block_mkperm(void *stream, const uint32_t *values, uint32_t size,
                                           uint32_t block_size, uint32_t bucket_count,
                                           uint32_t *perm, uint32_t *offsets,
                                           uint32_t *unique_out);

/* CUDAThreadState::poke, src/cuda_ts.cpp:988-1006: store `size` in {1,2,4,8} bytes from host `src`. */
DRJIT_B200_API int drjit_b200_poke(void *stream, void *dst, const void *src, uint32_t size);

/* CUDAThreadState::aggregate, src/cuda_ts.cpp:1008-1026 (kernel resources/misc.cuh:41-61).
 * `agg` is a device(-accessible) array of `size` entries. */
DRJIT_B200_API int drjit_b200_aggregate(void *stream, void *dst,
                                        const struct drjit_b200_aggregation_entry *agg,
                                        uint32_t size);

/* Hand-written counterpart of the JIT-emitted scatter-reduce template,
 * jitc_cuda_render_scatter_reduce(), src/cuda_scatter.cpp:246-354 (dr.scatter_reduce):
 * target[index[i]] op= value[i] for i < size where mask[i] != 0 (mask may be NULL).
 * mode: Auto/Local = warp-level pre-reduction of equal indices + privatised shared-memory
 * bins when the target fits; Direct = one atomic per element. Supported (op, type) pairs
 * follow the reference's capability table (src/op.cpp:2735-2822): Add: i32,u32,i64,u64,f16,f32,f64;
 * Min/Max: i32,u32,i64,u64,f16,f32,f64; And/Or: i32,u32,i64,u64. Others: EUNSUPPORTED.
 * f16 uses the two-wide f16 reductions with an identity partner (src/cuda_scatter.cpp:291-332):
 * like the reference, the 4-byte word holding the last element of an odd-sized target is
 * touched in full. */
DRJIT_B200_API int drjit_b200_scatter_reduce(void *stream, int vt, int op, int mode, void *target,
                                             uint32_t target_size, const void *value,
                                             const uint32_t *index, const uint8_t *mask,
                                             uint32_t size);

/* ---- sharded / asynchronous forms (new: the reference is single-device) ---
 * Building blocks for one-process-per-GPU sharding: the shard-local
 * pass of each primitive with the cross-shard term supplied as a device scalar, so
 * that the only inter-GPU traffic is the tiny combine message. */

/* Full-array prefix reduction of a shard (block_size == size semantics) whose running
 * value starts at *carry_in (device scalar of type vt, NULL = identity). If total_out
 * (device scalar) is non-NULL it receives op(carry_in, reduction of the shard). */
DRJIT_B200_API int drjit_b200_prefix_reduce_carry(void *stream, int vt, int op, uint32_t size,
                                                  int exclusive, int reverse, const void *in,
                                                  void *out, const void *carry_in,
                                                  void *total_out);

/* Asynchronous compress: indices are offset by index_base (the shard's first global
 * element index); the count is written to the device scalar *count_dev. No sync. */
DRJIT_B200_API int drjit_b200_compress_async(void *stream, const uint8_t *in, uint32_t size,
                                             uint32_t index_base, uint32_t *out,
                                             uint32_t *count_dev);

/* Shard-local mkperm (single sorting group, asynchronous, no host table): `perm` receives
 * the shard's permutation with entries index_base + local index, and hist_dev[bucket_count]
 * (device, may be NULL) the shard's per-bucket key counts. The caller all-reduces the
 * histograms to obtain the global bucket table; each shard's slice of bucket b starts at the
 * exclusive scan of its own histogram. */
DRJIT_B200_API int drjit_b200_mkperm_sharded(void *stream, const uint32_t *values, uint32_t size,
                                             uint32_t bucket_count, uint32_t index_base,
                                             uint32_t *perm, uint32_t *hist_dev);

/* Fill device arrays with the synthetic inputs of the benchmark (fmix32 of the element
 * index, ext/drjit-core/tests/reductions.cpp:5-13) without a host round trip.
 * kind: 0 = u32 fmix32((start+i)^xor_) & and_; 1 = f32 unit float (top 24 bits / 2^24);
 *       2 = u8 mask ((fmix32(start+i) & 0xff) < and_). */
DRJIT_B200_API int drjit_b200_fill_fmix32(void *stream, int kind, void *out, uint64_t start,
                                          uint64_t n, uint32_t xor_, uint32_t and_);

#if defined(__cplusplus)
}
#endif

#endif /* DRJIT_B200_H */
