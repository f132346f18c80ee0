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

/* KernelType, include/drjit-core/jit.h:2597-2632: the tag a launch carries in the kernel
 * history. Values 0..11 are the reference's; >= 256 are primitives the reference has no
 * precompiled kernel for (they run inside JIT kernels there). */
enum drjit_b200_kernel_type {
    DRJIT_B200_KT_JIT = 0, DRJIT_B200_KT_BLOCK_REDUCE = 1, DRJIT_B200_KT_BLOCK_PREFIX_REDUCE = 2,
    DRJIT_B200_KT_DOT = 3, DRJIT_B200_KT_BATCHED_GEMM = 4, DRJIT_B200_KT_COMPRESS = 5,
    DRJIT_B200_KT_MKPERM = 6, DRJIT_B200_KT_MEMCPY = 7, DRJIT_B200_KT_MEMSET = 8,
    DRJIT_B200_KT_POKE = 9, DRJIT_B200_KT_AGGREGATE = 10, DRJIT_B200_KT_LLVM_HOST_FUNC = 11,
    DRJIT_B200_KT_SCATTER_REDUCE = 256, DRJIT_B200_KT_SORT = 257, DRJIT_B200_KT_PEER_EXCHANGE = 258
};

/* Subset of JitFlag (jit.h:1680-1783) that acts at the seam (src/cuda_ts.cpp:19-46). */
enum drjit_b200_flag {
    DRJIT_B200_FLAG_KERNEL_HISTORY = 1,   /* JitFlag::KernelHistory: record an entry per primitive call */
    DRJIT_B200_FLAG_LAUNCH_BLOCKING = 2   /* JitFlag::LaunchBlocking: synchronise the stream after each call */
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

/* CallBucket, include/drjit-core/jit.h:2470-2480 (16 bytes). jit_var_call_reduce overwrites the rows of
 * the bucket table IN PLACE with this structure (src/call.cpp:1358-1378): a table row
 * {id, start, size, 0} written by drjit_b200_block_mkperm / drjit_b200_call_reduce has the same size, so
 * the pinned `offsets` buffer doubles as the CallBucket array the dispatcher returns. `index` is a
 * tracer variable id (the sub-range of `perm`), filled in above the seam. */
struct drjit_b200_call_bucket {
    void *ptr;               /* resolved instance pointer (registry) */
    uint32_t index;          /* variable holding perm[start .. start + size) */
    uint32_t id;             /* original instance ID */
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
 * (counterpart of jitc_cuda_shutdown(), src/cuda_core.cpp:541-582). The caller must make sure
 * that no other thread is inside the library and that all streams it was used on are idle. */
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

/* Scratch memory hooks, continued: `free_fn` MUST be stream-ordered like jitc_free()
 * (src/malloc.cpp:202-260: the block is recycled only after the work enqueued before the free
 * has finished): the library releases its temporaries as soon as the host call returns, while
 * the kernels using them are still in flight on `stream`. Both hooks may be called from any
 * thread that calls into the library. Install them before the first primitive call. */

/* Make sure the library's own arena for `stream` (on the current device) holds at least `bytes`
 * without a later cudaMalloc, e.g. before capturing calls into a CUDA graph: inside a capture the
 * arena cannot grow and a primitive that needs more fails with DRJIT_B200_ECUDA. Capturable entry
 * points are the asynchronous ones (everything except compress / all / any / block_mkperm with a
 * host table, which block on the stream like their reference counterparts). A warm-up call of the
 * same primitive at the same size has the same effect. No-op when allocator hooks are installed. */
DRJIT_B200_API int drjit_b200_reserve_scratch(void *stream, size_t bytes);

/* Number of kernels launched by this library on the calling thread since the last
 * call with reset != 0 (KernelHistory-style accounting, src/cuda_ts.cpp:19-46). */
DRJIT_B200_API uint64_t drjit_b200_launch_count(int reset);

/* ---- kernel history / launch blocking at the seam (src/cuda_ts.cpp:19-46) ----
 * The reference's submit_gpu() brackets every primitive launch with two CUDA events and appends
 * a KernelHistoryEntry{backend, type, size, ...} when JitFlag::KernelHistory is set, and
 * synchronises after the launch when JitFlag::LaunchBlocking is set. Two ways to keep that
 * working behind the seam:
 *
 * (a) Launch hook (what the drjit-core adapter installs, drjit_b200_thread_state.h): called on
 *     the calling thread with phase 0 before the first launch of a primitive call and with
 *     phase 1 after its last launch. `*cookie` is NULL at phase 0 and carries whatever the hook
 *     stores there over to phase 1. `launches` (phase 1 only) is the number of kernels the call
 *     enqueued -- 0 means the call was a no-op (size 0, plain memcpy ...) and nothing should be
 *     recorded. A primitive is one entry (the reference records one per launch; full reductions
 *     and scans are single launches here, mkperm is four). */
typedef void (*drjit_b200_launch_hook)(void *user, int phase, int kernel_type, uint32_t size,
                                       void *stream, uint32_t launches, void **cookie);
DRJIT_B200_API int drjit_b200_set_launch_hook(drjit_b200_launch_hook hook, void *user);

/* (b) Built-in history for use without Dr.Jit: flags are thread-local like JitFlag. */
DRJIT_B200_API int drjit_b200_set_flags(uint32_t flags);
DRJIT_B200_API uint32_t drjit_b200_flags(void);

struct drjit_b200_history_entry {       /* the fields of KernelHistoryEntry this path fills in */
    uint32_t type;                      /* enum drjit_b200_kernel_type */
    uint32_t size;                      /* number of array entries processed */
    uint32_t launches;                  /* kernels enqueued by the call */
    float execution_time;               /* ms between the bracketing events */
};
/* Copies up to `max_entries` recorded entries of the calling thread (oldest first) into
 * `entries`, waits for their events to obtain the timings, clears the history and returns the
 * number copied (jit_kernel_history(), jit.h:2700-2710). */
DRJIT_B200_API uint32_t drjit_b200_kernel_history(struct drjit_b200_history_entry *entries,
                                                  uint32_t max_entries);
DRJIT_B200_API void drjit_b200_kernel_history_clear(void);

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
 * bucket appear in unspecified order, like in the reference's "small" / "large" variants.
 * Keys >= bucket_count are undefined behaviour in the reference; here every path counts and
 * places them in the LAST bucket (no out-of-bounds access, same result at every input size). */
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

/* Packet form of the scatter-reduce template, jitc_cuda_render_scatter_reduce_packet(),
 * src/cuda_packet.cpp:168-327 (jit_var_scatter_packet, jit.h:1107-1120; dr.scatter_reduce /
 * dr.scatter_add of an ArrayNf, e.g. film accumulation):
 *   target[index[i] * count + k] op= values[k][i]   for k < count, i < size where mask[i] != 0.
 * `values` is a HOST array of `count` device pointers, one contiguous array of `size` elements per
 * component (the layout of a Dr.Jit array of JIT arrays); `target` holds `target_packets` packets of
 * `count` consecutive elements. `count` must be even (cuda_packet.cpp:184-186) and <= 16: EINVAL
 * otherwise. (type, op) pairs as for drjit_b200_scatter_reduce. f32 Add and f16 Add/Min/Max leave as
 * vector reductions of up to 16 bytes (red.global.v4.f32.add, red.global.v8.f16.<op>.noftz,
 * cuda_packet.cpp:224-259) when `target` is aligned to the vector; all other pairs as one reduction
 * per component. mode Local pre-combines lanes with equal index (not for f16, as in the reference,
 * cuda_packet.cpp:200-207); Auto = Direct. */
DRJIT_B200_API int drjit_b200_scatter_reduce_packet(void *stream, int vt, int op, int mode, void *target,
                                                    uint32_t target_packets, const void *const *values,
                                                    uint32_t count, const uint32_t *index,
                                                    const uint8_t *mask, uint32_t size);

/* Standalone form of jitc_cuda_render_scatter_inc(), src/cuda_scatter.cpp:356-393
 * (jit_var_scatter_inc, jit.h:1126-1143; dr.scatter_inc): for every i < size with mask[i] != 0
 * (mask may be NULL) atomically  out[i] = target[index[i]]++ ;  out[i] = 0 for masked elements
 * (:361-364). `index` may be NULL: every element increments counter 0 (the queue form,
 * dr.scatter_inc(counter, 0)). Which element receives which of the slots handed out for one counter is
 * unspecified (the reference aggregates per warp, here small counter arrays are aggregated per CTA);
 * the slots of a counter are distinct and contiguous from its value before the call. Indices
 * >= target_size (undefined behaviour in the reference) are ignored and yield 0. */
DRJIT_B200_API int drjit_b200_scatter_inc(void *stream, uint32_t *target, uint32_t target_size,
                                          const uint32_t *index, const uint8_t *mask, uint32_t size,
                                          uint32_t *out);

/* jit_var_call_reduce, src/call.cpp:1268-1389 (dr.dispatch / vcall reordering), as one call:
 * block_mkperm of the callable IDs with block_size == size, plus
 *  - the table of non-empty buckets already in the order the dispatcher uses -- decreasing bucket
 *    size (:1346-1356; the reference std::sorts the pinned table on the host after the wait; here the
 *    bucket-scan kernel sorts it in shared memory; ties in ascending id order), and
 *  - up to 4 argument arrays of 32-bit elements permuted on the way: payload_out[k][j] =
 *    payload_in[k][perm[j]] for all j, written by the scatter pass where `perm` is written (the
 *    reference gathers every argument through `perm` in a separate pass, src/extra/call.cpp:320-452).
 * offsets: host-pinned, 4*bucket_count+1 entries (required layout as in drjit_b200_block_mkperm).
 * Waits for the table like drjit_b200_block_mkperm; perm / payload_out are complete in stream order. */
DRJIT_B200_API int drjit_b200_call_reduce(void *stream, const uint32_t *ids, uint32_t size, uint32_t bucket_count,
                                          uint32_t *perm, uint32_t *offsets, uint32_t n_payloads,
                                          const void *const *payload_in, void *const *payload_out,
                                          uint32_t *unique_out);

/* dr.sort / dr.argsort, drjit/__init__.py:1698-1772 (`_radix_sort`: on the GPU four (eight for 64-bit
 * types) 8-bit LSD passes, each a digit kernel + jit_block_mkperm + one gather per carried array).
 * Here every pass moves the keys and the index payload itself (histogram launch + offsets launch +
 * stable scatter launch, 20 bytes per element and pass instead of ~44; no permutation array, no
 * gathers); the order-preserving transform of _to_ordinal_32/_64 (:1483-1520) is applied on the fly.
 * keys: `size` entries of type vt in {i32, u32, f32, i64, u64, f64}; keys_out (same type, may be NULL):
 * the sorted keys; index_out (u32, may be NULL): the stable sorting permutation (dr.argsort), i.e.
 * keys_out[i] == keys[index_out[i]], equal keys in input order. Float order is that of the reference's
 * ordinal transform: -NaN < -inf < ... < -0.0 < +0.0 < ... < +inf < +NaN. keys_out must not alias
 * keys. Asynchronous. Sorting of independent groups (block_size < size) is not implemented here:
 * compose drjit_b200_block_mkperm passes as the reference does. */
DRJIT_B200_API int drjit_b200_sort(void *stream, int vt, uint32_t size, int descending, const void *keys,
                                   void *keys_out, uint32_t *index_out);

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

/* ---- multi-GPU: primitives fused with their combine step over peer memory -------------------
 * (new: the reference is single-device). One communicator per GPU of one
 * NVSwitch box (<= DRJIT_B200_COMM_MAX_RANKS ranks). Each rank owns a "window" of device memory
 * that all peers map (CUDA IPC between processes, plain peer access inside one process); the
 * combine step of every primitive below is a few stores + flag spins over NVLink inside the
 * kernel that produces the partial -- no library collective, no second launch, no host round
 * trip. All ranks must issue the same sequence of comm calls, each rank on ONE stream; results
 * that are defined for all ranks are bit-identical on all of them (fixed rank-order folds).
 * A peer that never arrives makes the waiting kernel trap after ~20 s (loud failure, no hang). */
#define DRJIT_B200_COMM_MAX_RANKS 8
#define DRJIT_B200_COMM_HANDLE_BYTES 64
#define DRJIT_B200_COMM_MAX_BUCKETS 16384      /* histogram exchange of drjit_b200_comm_mkperm */

/* Creates this rank's communicator on the current device. bulk_bytes: capacity of the all-reduce
 * staging area (>= padded byte size of the largest array passed to drjit_b200_comm_allreduce;
 * 0 if unused). The window (1 MiB + 2 * bulk_bytes) is allocated and zeroed here. */
DRJIT_B200_API int drjit_b200_comm_create(uint32_t rank, uint32_t world, size_t bulk_bytes, void **comm_out);
/* One process per GPU: write this rank's DRJIT_B200_COMM_HANDLE_BYTES-byte window handle to
 * handle_out, all-gather the handles by any means (the launcher's process group, MPI, a file), then connect
 * with the `world` handles in rank order. */
DRJIT_B200_API int drjit_b200_comm_handle(void *comm, void *handle_out);
DRJIT_B200_API int drjit_b200_comm_connect(void *comm, const void *handles);
/* All ranks inside one process (one thread or several; how Dr.Jit itself drives several devices,
 * src/cuda_core.cpp:518-536): comms[r] = communicator of rank r. Enables peer access as needed. */
DRJIT_B200_API int drjit_b200_comm_connect_local(void **comms, uint32_t world);
/* Every rank must have finished its last comm call (and synchronised) before any rank destroys. */
DRJIT_B200_API int drjit_b200_comm_destroy(void *comm);

/* fold: which ranks' partials the result combines */
enum drjit_b200_comm_fold {
    DRJIT_B200_FOLD_ALL = 0,      /* all ranks: the reduction of the global array */
    DRJIT_B200_FOLD_LOWER = 1,    /* ranks below the caller: carry of a forward scan */
    DRJIT_B200_FOLD_HIGHER = 2    /* ranks above the caller: carry of a reverse scan */
};

/* jit_reduce / dr.sum|prod|min|max over a sharded array: reduction of this rank's shard
 * in[0..size) and the fold over the ranks in ONE launch; out[0] (device) receives the result.
 * size == 0 (empty trailing shard) contributes the identity. Same (vt, op) table as
 * drjit_b200_block_reduce. */
DRJIT_B200_API int drjit_b200_comm_reduce(void *comm, void *stream, int vt, int op, int fold, uint32_t size,
                                          const void *in, void *out);
/* jit_reduce_dot over sharded arrays, one launch. */
DRJIT_B200_API int drjit_b200_comm_reduce_dot(void *comm, void *stream, int vt, const void *a, const void *b,
                                              uint32_t size, void *out);
/* dr.all / dr.any over a sharded mask (synchronous, like jitc_all / jitc_any). */
DRJIT_B200_API int drjit_b200_comm_all(void *comm, void *stream, const uint8_t *values, uint32_t size, int *result);
DRJIT_B200_API int drjit_b200_comm_any(void *comm, void *stream, const uint8_t *values, uint32_t size, int *result);
/* Prefix reduction of a global array cut into contiguous shards in rank order.
 * materialise != 0: out[i] is the global prefix value (two launches: shard reduction with the totals
 *   exchanged inside its last CTA, then the single-pass scan seeded with the carry; 12 bytes per
 *   4-byte element instead of 8 -- a rank cannot emit before all lower ranks have read their shard).
 * materialise == 0: shard-offset form, out[i] = prefix inside the shard and *offset_out = fold over
 *   the lower (reverse: higher) ranks, global[i] = op(offset, out[i]); 8 bytes per element.
 * offset_out: device scalar of type vt (required for the offset form, optional otherwise). */
DRJIT_B200_API int drjit_b200_comm_prefix_reduce(void *comm, void *stream, int vt, int op, uint32_t size,
                                                 int exclusive, int reverse, const void *in, void *out,
                                                 void *offset_out, int materialise);
/* jit_compress of this rank's shard (indices are index_base + local index) and the exchange of the
 * per-rank counts: counts_host[0..world) (host). One launch: the exchange happens inside the
 * compaction kernel, which writes the W counts to pinned host memory. The call blocks until the counts
 * are there and returns them; `out` is complete in stream order (the contract jit_block_mkperm has for
 * `perm`, cuda_ts.cpp:953-967), so anything enqueued on `stream` afterwards sees the full list. */
DRJIT_B200_API int drjit_b200_comm_compress(void *comm, void *stream, const uint8_t *in, uint32_t size,
                                            uint32_t index_base, uint32_t *out, uint32_t *counts_host);
/* jit_block_mkperm of one sorting group sharded over the ranks: perm (device, `size` entries) = this
 * shard's permutation with entries index_base + local index; hist_dev[bucket_count] (device, may be
 * NULL) = shard counts; rank_base_dev[bucket_count] (device, may be NULL) = start of this rank's keys
 * of every bucket in the global rank-major (stable) order; offsets (host-pinned, 4*bucket_count+1,
 * may be NULL) = table of non-empty buckets of the GLOBAL array {id, start, size, 0} + unique count,
 * identical on every rank; the call waits for the table like drjit_b200_block_mkperm.
 * bucket_count <= DRJIT_B200_COMM_MAX_BUCKETS. The histograms are exchanged inside the bucket-scan
 * kernel. */
DRJIT_B200_API int drjit_b200_comm_mkperm(void *comm, void *stream, const uint32_t *values, uint32_t size,
                                          uint32_t bucket_count, uint32_t index_base, uint32_t *perm,
                                          uint32_t *hist_dev, uint32_t *rank_base_dev, uint32_t *offsets,
                                          uint32_t *unique_out);
/* In-place sum of `count` elements of data (device, 16-byte aligned; f32/f64/i32/u32/i64/u64) over all
 * ranks -- the bins of a sharded dr.scatter_reduce(Add). One kernel: reduce-scatter + all-gather
 * through the windows, every element folded once in rank order (bit-identical on all ranks). */
DRJIT_B200_API int drjit_b200_comm_allreduce(void *comm, void *stream, int vt, int op, void *data, uint32_t count);
/* Building blocks: all-gather of a small payload (bytes % 4 == 0, <= 64 KiB per rank; dst holds
 * world * bytes, device or device-mapped host memory) and fold of one scalar per rank. */
DRJIT_B200_API int drjit_b200_comm_allgather(void *comm, void *stream, const void *src, uint32_t bytes, void *dst);
DRJIT_B200_API int drjit_b200_comm_fold(void *comm, void *stream, int vt, int op, int fold, const void *src, void *dst);

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
