/*
 * oracle.c -- CPU restatement of Dr.Jit-Core's data-parallel primitives.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (drjit_b200/, include/) may
 * import, link or call this file. It is the checker used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * Parity status: PINNED. tests/test_oracle.py checks every function below against
 *   (a) the reference's own known-answer tests (ext/drjit-core/tests/reductions.cpp,
 *       tests/test_memop.py:734-756, tests/test_reduction.py:370-398), and
 *   (b) outputs of the UNMODIFIED reference libdrjit-core.so (LLVM backend CPU
 *       primitives) built by oracle/ref_build/Makefile into oracle/_ref/, both live
 *       (when oracle/_ref exists) and through committed fixtures in tests/golden/.
 *
 * Each function is a plain serial loop that follows the semantics of the
 * reference's LLVMThreadState implementation with a single worker (pool_size()==1),
 * i.e. chunk_size == block_size and one work unit. Citations are relative to
 * /root/reference/ext/drjit-core.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#define EXPORT __attribute__((visibility("default")))

/* VarType / ReduceOp numeric values: include/drjit-core/jit.h:597-611, :990-1014 */
enum { VT_BOOL = 1, VT_I8 = 3, VT_U8 = 4, VT_I16 = 5, VT_U16 = 6, VT_I32 = 7, VT_U32 = 8,
       VT_I64 = 9, VT_U64 = 10, VT_F16 = 13, VT_F32 = 14, VT_F64 = 15 };
enum { OP_ADD = 1, OP_MUL = 2, OP_MIN = 3, OP_MAX = 4, OP_AND = 5, OP_OR = 6 };

typedef _Float16 half;

/* ------------------------------------------------------------------------
 * Input generators. fmix32: tests/reductions.cpp:5-13 (note the `h += 1`).
 * ---------------------------------------------------------------------- */
static inline uint32_t fmix32(uint32_t h) {
    h += 1;
    h ^= h >> 16; h *= 0x85ebca6bu;
    h ^= h >> 13; h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
static inline uint64_t fmix32_64(uint64_t h) { /* same mixer instantiated for uint64_t (reductions.cpp:5, Value=uint64_t) */
    h += 1;
    h ^= h >> 16; h *= 0x85ebca6bull;
    h ^= h >> 13; h *= 0xc2b2ae35ull;
    h ^= h >> 16;
    return h;
}

EXPORT void oracle_fill_fmix32_u32(uint32_t *out, uint64_t start, uint64_t n, uint32_t xor_, uint32_t and_) {
    for (uint64_t i = 0; i < n; ++i)
        out[i] = fmix32((uint32_t) (start + i) ^ xor_) & and_;
}
EXPORT void oracle_fill_fmix32_u64(uint64_t *out, uint64_t start, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i)
        out[i] = fmix32_64(start + i);
}
/* unit floats: top 24 bits / 2^24 in [0,1)   */
EXPORT void oracle_fill_unit_f32(float *out, uint64_t start, uint64_t n, uint32_t xor_) {
    for (uint64_t i = 0; i < n; ++i)
        out[i] = (float) (fmix32((uint32_t) (start + i) ^ xor_) >> 8) * (1.0f / 16777216.0f);
}
/* mask[i] = (fmix32(i) & 0xff) < threshold */
EXPORT void oracle_fill_mask_u8(uint8_t *out, uint64_t start, uint64_t n, uint32_t threshold) {
    for (uint64_t i = 0; i < n; ++i)
        out[i] = (fmix32((uint32_t) (start + i)) & 0xffu) < threshold ? 1 : 0;
}

/* ------------------------------------------------------------------------
 * Reduction functors: src/llvm_red.h:10-85. `A` is the accumulator ("Value")
 * type: float for half (llvm_red.h:11), T otherwise.
 * Min/Max follow std::min/std::max (llvm_red.h:36,50): min(a,b) = b<a ? b : a.
 * ---------------------------------------------------------------------- */
#define RED_ADD(a, b) ((a) + (b))
#define RED_MUL(a, b) ((a) * (b))
#define RED_MIN(a, b) ((b) < (a) ? (b) : (a))
#define RED_MAX(a, b) ((a) < (b) ? (b) : (a))
#define RED_AND(a, b) ((a) & (b))
#define RED_OR(a, b)  ((a) | (b))

/* block_reduce kernel: llvm_red.h:87-118 with chunk_size == block_size */
#define DEF_BLOCK_REDUCE(NAME, T, A, INIT, RED)                                         \
    static void block_reduce_##NAME(uint32_t size, uint32_t bs, const void *in_,       \
                                    void *out_) {                                       \
        const T *in = (const T *) in_; T *out = (T *) out_;                             \
        uint32_t blocks = (uint32_t) (((uint64_t) size + bs - 1) / bs);                 \
        for (uint32_t b = 0; b < blocks; ++b) {                                         \
            uint64_t start = (uint64_t) b * bs, end = start + bs;                       \
            if (end > size) end = size;                                                 \
            A accum = (A) (INIT);                                                       \
            for (uint64_t j = start; j < end; ++j)                                      \
                accum = RED(accum, (A) in[j]);                                          \
            out[b] = (T) accum;                                                         \
        }                                                                               \
    }

/* block_prefix_reduce kernel: llvm_red.h:127-179 (scratch == NULL) */
#define DEF_BLOCK_PREFIX(NAME, T, A, INIT, RED)                                         \
    static void block_prefix_##NAME(uint32_t size, uint32_t bs, int exclusive,          \
                                    int reverse, const void *in_, void *out_) {         \
        const T *in = (const T *) in_; T *out = (T *) out_;                             \
        uint32_t blocks = (uint32_t) (((uint64_t) size + bs - 1) / bs);                 \
        for (uint32_t b = 0; b < blocks; ++b) {                                         \
            uint64_t start = (uint64_t) b * bs, end = start + bs;                       \
            if (end > size) end = size;                                                 \
            A accum = (A) (INIT);                                                       \
            if (!reverse) {                                                             \
                for (uint64_t j = start; j < end; ++j) {                                \
                    A value = (A) in[j], prev = accum;                                  \
                    accum = RED(accum, value);                                          \
                    out[j] = (T) (exclusive ? prev : accum);                            \
                }                                                                       \
            } else {                                                                    \
                for (uint64_t j = end; j > start; --j) {                                \
                    uint64_t k = j - 1;                                                 \
                    A value = (A) in[k], prev = accum;                                  \
                    accum = RED(accum, value);                                          \
                    out[k] = (T) (exclusive ? prev : accum);                            \
                }                                                                       \
            }                                                                           \
        }                                                                               \
    }

#define DEF_BOTH(NAME, T, A, INIT, RED) \
    DEF_BLOCK_REDUCE(NAME, T, A, INIT, RED) DEF_BLOCK_PREFIX(NAME, T, A, INIT, RED)

/* integer types: all six ops. identities: llvm_red.h:13,21,29-34,41-48,54,70 */
#define DEF_INT(TN, T, TMIN, TMAX)                      \
    DEF_BOTH(add_##TN, T, T, 0, RED_ADD)                \
    DEF_BOTH(mul_##TN, T, T, 1, RED_MUL)                \
    DEF_BOTH(min_##TN, T, T, TMAX, RED_MIN)             \
    DEF_BOTH(max_##TN, T, T, TMIN, RED_MAX)             \
    DEF_BOTH(and_##TN, T, T, (T) -1, RED_AND)           \
    DEF_BOTH(or_##TN,  T, T, 0, RED_OR)

DEF_INT(u8,  uint8_t,  0, UINT8_MAX)
DEF_INT(i32, int32_t,  INT32_MIN, INT32_MAX)
DEF_INT(u32, uint32_t, 0, UINT32_MAX)
DEF_INT(i64, int64_t,  INT64_MIN, INT64_MAX)
DEF_INT(u64, uint64_t, 0, UINT64_MAX)

/* floating point types: add/mul/min/max; And/Or return 0 (llvm_red.h:58-83).
 * `A` = float for half. The *_acc64 family accumulates f16/f32 in double: it is
 * the tolerance anchor for large float sums (a serial f32 accumulator is not). */
#define DEF_FLT(TN, T, A)                               \
    DEF_BOTH(add_##TN, T, A, 0, RED_ADD)                \
    DEF_BOTH(mul_##TN, T, A, 1, RED_MUL)                \
    DEF_BOTH(min_##TN, T, A, INFINITY, RED_MIN)         \
    DEF_BOTH(max_##TN, T, A, -INFINITY, RED_MAX)

DEF_FLT(f16, half, float)
DEF_FLT(f32, float, float)
DEF_FLT(f64, double, double)
DEF_FLT(f16_acc64, half, double)
DEF_FLT(f32_acc64, float, double)

typedef void (*block_reduce_fn)(uint32_t, uint32_t, const void *, void *);
typedef void (*block_prefix_fn)(uint32_t, uint32_t, int, int, const void *, void *);

#define PICK_INT(KIND, TN)                              \
    switch (op) {                                       \
        case OP_ADD: return KIND##_add_##TN;            \
        case OP_MUL: return KIND##_mul_##TN;            \
        case OP_MIN: return KIND##_min_##TN;            \
        case OP_MAX: return KIND##_max_##TN;            \
        case OP_AND: return KIND##_and_##TN;            \
        case OP_OR:  return KIND##_or_##TN;             \
        default: return NULL;                           \
    }
#define PICK_FLT(KIND, TN)                              \
    switch (op) {                                       \
        case OP_ADD: return KIND##_add_##TN;            \
        case OP_MUL: return KIND##_mul_##TN;            \
        case OP_MIN: return KIND##_min_##TN;            \
        case OP_MAX: return KIND##_max_##TN;            \
        default: return NULL;                           \
    }

/* type dispatch: llvm_red.h:181-222 (UInt8, Int32, UInt32, Int64, UInt64, f16, f32, f64) */
static block_reduce_fn pick_block_reduce(int vt, int op, int acc64) {
    switch (vt) {
        case VT_U8:  PICK_INT(block_reduce, u8)
        case VT_I32: PICK_INT(block_reduce, i32)
        case VT_U32: PICK_INT(block_reduce, u32)
        case VT_I64: PICK_INT(block_reduce, i64)
        case VT_U64: PICK_INT(block_reduce, u64)
        case VT_F16: if (acc64) { PICK_FLT(block_reduce, f16_acc64) } else { PICK_FLT(block_reduce, f16) }
        case VT_F32: if (acc64) { PICK_FLT(block_reduce, f32_acc64) } else { PICK_FLT(block_reduce, f32) }
        case VT_F64: PICK_FLT(block_reduce, f64)
        default: return NULL;
    }
}
static block_prefix_fn pick_block_prefix(int vt, int op, int acc64) {
    switch (vt) {
        case VT_U8:  PICK_INT(block_prefix, u8)
        case VT_I32: PICK_INT(block_prefix, i32)
        case VT_U32: PICK_INT(block_prefix, u32)
        case VT_I64: PICK_INT(block_prefix, i64)
        case VT_U64: PICK_INT(block_prefix, u64)
        case VT_F16: if (acc64) { PICK_FLT(block_prefix, f16_acc64) } else { PICK_FLT(block_prefix, f16) }
        case VT_F32: if (acc64) { PICK_FLT(block_prefix, f32_acc64) } else { PICK_FLT(block_prefix, f32) }
        case VT_F64: PICK_FLT(block_prefix, f64)
        default: return NULL;
    }
}

static size_t type_size(int vt) {
    switch (vt) {
        case VT_BOOL: case VT_I8: case VT_U8: return 1;
        case VT_I16: case VT_U16: case VT_F16: return 2;
        case VT_I32: case VT_U32: case VT_F32: return 4;
        case VT_I64: case VT_U64: case VT_F64: return 8;
        default: return 0;
    }
}

/* reduction identity as raw bits: src/var.cpp:2642-2652 */
EXPORT uint64_t oracle_reduce_identity(int vt, int op) {
    uint64_t r = 0;
    switch (op) {
        case OP_ADD: case OP_OR: return 0;
        case OP_AND: return type_size(vt) == 8 ? ~0ull : ((1ull << (8 * type_size(vt))) - 1);
        case OP_MUL:
            switch (vt) {
                case VT_F16: { half h = 1; memcpy(&r, &h, 2); return r; }
                case VT_F32: { float f = 1; memcpy(&r, &f, 4); return r; }
                case VT_F64: { double d = 1; memcpy(&r, &d, 8); return r; }
                default: return 1;
            }
        case OP_MIN: /* largest value */
            switch (vt) {
                case VT_U8: return UINT8_MAX; case VT_I8: return INT8_MAX;
                case VT_U16: return UINT16_MAX; case VT_I16: return INT16_MAX;
                case VT_U32: return UINT32_MAX; case VT_I32: return INT32_MAX;
                case VT_U64: return UINT64_MAX; case VT_I64: return INT64_MAX;
                case VT_F16: return 0x7C00; case VT_F32: return 0x7F800000u;
                case VT_F64: return 0x7FF0000000000000ull;
                default: return 0;
            }
        case OP_MAX: /* smallest value */
            switch (vt) {
                case VT_U8: case VT_U16: case VT_U32: case VT_U64: return 0;
                case VT_I8: return (uint8_t) INT8_MIN; case VT_I16: return (uint16_t) INT16_MIN;
                case VT_I32: return (uint32_t) INT32_MIN; case VT_I64: return (uint64_t) INT64_MIN;
                case VT_F16: return 0xFC00; case VT_F32: return 0xFF800000u;
                case VT_F64: return 0xFFF0000000000000ull;
                default: return 0;
            }
    }
    return r;
}

/* LLVMThreadState::block_reduce, src/llvm_ts.cpp:265-350.
 * returns 0 on success, -1 = "invalid block size" (jitc_raise, :271-274),
 * -2 = unsupported type/op (llvm_red.h:195,219). size==0 is a no-op (:269). */
EXPORT int oracle_block_reduce(int vt, int op, uint32_t size, uint32_t block_size,
                               const void *in, void *out, int acc64) {
    if (size == 0) return 0;
    if (block_size == 0 || block_size > size) return -1;
    if (block_size == 1) { memcpy(out, in, (size_t) size * type_size(vt)); return 0; } /* :275-278 */
    block_reduce_fn f = pick_block_reduce(vt, op, acc64);
    if (!f) return -2;
    f(size, block_size, in, out);
    return 0;
}

/* LLVMThreadState::block_prefix_reduce, src/llvm_ts.cpp:352-460. In-place allowed. */
EXPORT int oracle_block_prefix_reduce(int vt, int op, uint32_t size, uint32_t block_size,
                                      int exclusive, int reverse, const void *in, void *out,
                                      int acc64) {
    if (size == 0) return 0;
    if (block_size == 0 || block_size > size) return -1;
    size_t ts = type_size(vt);
    if (block_size == 1) { /* :364-372 */
        if (exclusive) {
            uint64_t ident = oracle_reduce_identity(vt, op);
            for (uint32_t i = 0; i < size; ++i) memcpy((uint8_t *) out + i * ts, &ident, ts);
        } else if (in != out) {
            memcpy(out, in, (size_t) size * ts);
        }
        return 0;
    }
    block_prefix_fn f = pick_block_prefix(vt, op, acc64);
    if (!f) return -2;
    f(size, block_size, exclusive, reverse, in, out);
    return 0;
}

/* reduce_dot: src/llvm_red.h:227-240 (serial std::fma chain), llvm_ts.cpp:479-516.
 * acc64: accumulate in double without fma contraction issues (tolerance anchor). */
EXPORT int oracle_reduce_dot(int vt, const void *a_, const void *b_, uint32_t size, void *out, int acc64) {
    switch (vt) {
        case VT_F16: {
            const half *a = a_, *b = b_;
            if (acc64) { double r = 0; for (uint32_t i = 0; i < size; ++i) r += (double) a[i] * (double) b[i]; *(half *) out = (half) r; }
            else { half r = 0; for (uint32_t i = 0; i < size; ++i) r = (half) fmaf((float) a[i], (float) b[i], (float) r); *(half *) out = r; }
            return 0;
        }
        case VT_F32: {
            const float *a = a_, *b = b_;
            if (acc64) { double r = 0; for (uint32_t i = 0; i < size; ++i) r += (double) a[i] * (double) b[i]; *(float *) out = (float) r; }
            else { float r = 0; for (uint32_t i = 0; i < size; ++i) r = fmaf(a[i], b[i], r); *(float *) out = r; }
            return 0;
        }
        case VT_F64: {
            const double *a = a_, *b = b_;
            double r = 0; for (uint32_t i = 0; i < size; ++i) r = fma(a[i], b[i], r); *(double *) out = r;
            return 0;
        }
        default: return -2;
    }
}

/* compress: src/llvm_ts.cpp:706-780 (blocks == 1). Note `accum += value`: the
 * mask bytes are added, entries are required to be 0/1 (jit.h:2375-2380). */
EXPORT uint32_t oracle_compress(const uint8_t *in, uint32_t size, uint32_t *out) {
    uint32_t accum = 0;
    for (uint32_t i = 0; i < size; ++i) {
        uint32_t value = in[i];
        if (value) out[accum] = i;
        accum += value;
    }
    return accum;
}

/* block_mkperm: src/llvm_ts.cpp:785-933 (stable counting sort per group of
 * `block_size` keys; offsets table {id,start,size,0} in ascending bucket order and
 * unique count only when there is a single group, :861-877).
 * returns unique count (0 when offsets == NULL or n_blocks > 1), -1 if bucket_count == 0
 * (jitc_fail in the reference, :791-792), -3 if a key is out of range. */
EXPORT int64_t oracle_block_mkperm(const uint32_t *ptr, uint32_t size, uint32_t block_size,
                                   uint32_t bucket_count, uint32_t *perm, uint32_t *offsets) {
    if (size == 0) return 0;
    if (bucket_count == 0) return -1;
    uint32_t n_blocks = (uint32_t) (((uint64_t) size + block_size - 1) / block_size);
    uint32_t *buckets = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) bucket_count);
    uint32_t unique = 0;
    for (uint32_t g = 0; g < n_blocks; ++g) {
        uint64_t start = (uint64_t) g * block_size, end = start + block_size;
        if (end > size) end = size;
        memset(buckets, 0, sizeof(uint32_t) * (size_t) bucket_count);
        for (uint64_t i = start; i < end; ++i) {
            if (ptr[i] >= bucket_count) { free(buckets); return -3; }
            buckets[ptr[i]]++;
        }
        uint32_t group_offset = 0;
        for (uint32_t b = 0; b < bucket_count; ++b) {
            uint32_t n = buckets[b];
            buckets[b] = (uint32_t) start + group_offset;
            if (n_blocks == 1 && n > 0 && offsets) {
                offsets[unique * 4 + 0] = b;
                offsets[unique * 4 + 1] = group_offset;
                offsets[unique * 4 + 2] = n;
                offsets[unique * 4 + 3] = 0;
                unique++;
            }
            group_offset += n;
        }
        for (uint64_t i = start; i < end; ++i)
            perm[buckets[ptr[i]]++] = (uint32_t) i;
    }
    free(buckets);
    if (offsets && n_blocks == 1) offsets[4 * (size_t) bucket_count] = unique; /* cuda_ts.cpp:948-951,974 */
    return (offsets && n_blocks == 1) ? unique : 0;
}

/* all/any: src/init.cpp:919-939 + src/util.cpp:153-211 (pad to a multiple of four
 * with the identity, And/Or-reduce as u32, combine the four bytes). */
EXPORT int oracle_all(const uint8_t *v, uint32_t size) {
    uint8_t r = 1; for (uint32_t i = 0; i < size; ++i) r &= v[i]; return r != 0;
}
EXPORT int oracle_any(const uint8_t *v, uint32_t size) {
    uint8_t r = 0; for (uint32_t i = 0; i < size; ++i) r |= v[i]; return r != 0;
}

/* scatter_reduce: target[index[i]] op= value[i], serial in index order
 * (semantics of jit_var_scatter with ReduceOp, jit.h:1076-1105; the CUDA path is
 * an atomic per element, src/cuda_scatter.cpp:246-354). mask may be NULL.
 * acc64: for f32 Add, accumulate into a double shadow array (tolerance anchor). */
#define SCATTER_LOOP(T, EXPR)                                                          \
    { T *t = (T *) target; const T *v = (const T *) value;                             \
      for (uint32_t i = 0; i < size; ++i) {                                            \
          if (mask && !mask[i]) continue;                                              \
          uint32_t k = index[i]; if (k >= target_size) return -3;                      \
          T a = t[k], b = v[i]; t[k] = (T) (EXPR); }                                   \
      return 0; }
#define SCATTER_INT(T)                                                                 \
    switch (op) {                                                                      \
        case OP_ADD: SCATTER_LOOP(T, a + b)                                            \
        case OP_MUL: SCATTER_LOOP(T, a * b)                                            \
        case OP_MIN: SCATTER_LOOP(T, RED_MIN(a, b))                                    \
        case OP_MAX: SCATTER_LOOP(T, RED_MAX(a, b))                                    \
        case OP_AND: SCATTER_LOOP(T, a & b)                                            \
        case OP_OR:  SCATTER_LOOP(T, a | b)                                            \
        default: return -2; }
#define SCATTER_FLT(T)                                                                 \
    switch (op) {                                                                      \
        case OP_ADD: SCATTER_LOOP(T, a + b)                                            \
        case OP_MUL: SCATTER_LOOP(T, a * b)                                            \
        case OP_MIN: SCATTER_LOOP(T, RED_MIN(a, b))                                    \
        case OP_MAX: SCATTER_LOOP(T, RED_MAX(a, b))                                    \
        default: return -2; }

EXPORT int oracle_scatter_reduce(int vt, int op, void *target, uint32_t target_size,
                                 const void *value, const uint32_t *index,
                                 const uint8_t *mask, uint32_t size, int acc64) {
    if (acc64 && vt == VT_F32 && op == OP_ADD) {
        double *shadow = (double *) malloc(sizeof(double) * (size_t) target_size);
        float *t = (float *) target; const float *v = (const float *) value;
        for (uint32_t k = 0; k < target_size; ++k) shadow[k] = t[k];
        for (uint32_t i = 0; i < size; ++i) {
            if (mask && !mask[i]) continue;
            uint32_t k = index[i]; if (k >= target_size) { free(shadow); return -3; }
            shadow[k] += v[i];
        }
        for (uint32_t k = 0; k < target_size; ++k) t[k] = (float) shadow[k];
        free(shadow);
        return 0;
    }
    switch (vt) {
        case VT_I32: SCATTER_INT(int32_t)
        case VT_U32: SCATTER_INT(uint32_t)
        case VT_I64: SCATTER_INT(int64_t)
        case VT_U64: SCATTER_INT(uint64_t)
        case VT_F16: SCATTER_FLT(half)
        case VT_F32: SCATTER_FLT(float)
        case VT_F64: SCATTER_FLT(double)
        default: return -2;
    }
}

/* scatter_inc: out[i] = target[index[i]]++, serial in index order (jit_var_scatter_inc, jit.h:1126-1143;
 * CUDA template src/cuda_scatter.cpp:356-393: masked lanes return 0, :361-364). index == NULL: every
 * element increments counter 0. The GPU assigns the slots of a counter in an unspecified order, so this
 * serial order is ONE valid result; tests compare the counters exactly and the slots per counter as sets
 * (the property the reference's own test checks, tests/test_memop.py:293-316). */
EXPORT int oracle_scatter_inc(uint32_t *target, uint32_t target_size, const uint32_t *index,
                              const uint8_t *mask, uint32_t size, uint32_t *out) {
    for (uint32_t i = 0; i < size; ++i) {
        out[i] = 0;
        if (mask && !mask[i]) continue;
        uint32_t k = index ? index[i] : 0;
        if (k >= target_size) return -3;
        out[i] = target[k]++;
    }
    return 0;
}

/* scatter_add in ReduceMode::Expand, the mode the reference's LLVM backend picks for targets of
 * up to 1 M entries (src/op.cpp:2845-2848, src/api.cpp:2073): the target is replicated once per
 * worker (jitc_var_expand, src/var.cpp:2866-2931; replication factor = pool size,
 * src/llvm_core.cpp:556-563), every worker scatters the 16384-element work units it picks up
 * (src/llvm_core.cpp:67) into its private copy without atomics, and the copies are folded into
 * copy 0 in worker order (reduce_expanded_impl, src/llvm_ts.cpp:1012-1028). The reference JIT-
 * compiles the scatter loop; this is the same loop in C. Used only by bench.py's CPU baseline
 * (the stub libLLVM of oracle/ref_build cannot JIT). Work units are dealt round-robin. */
#include <pthread.h>
struct expand_job { float *copies; const float *value; const uint32_t *index; uint32_t target_size, size, worker, workers; };
static void *expand_scatter_worker(void *arg) {
    struct expand_job *j = (struct expand_job *) arg;
    float *t = j->copies + (size_t) j->worker * j->target_size;
    const uint32_t unit = 16384;
    for (uint64_t lo = (uint64_t) j->worker * unit; lo < j->size; lo += (uint64_t) j->workers * unit) {
        uint64_t hi = lo + unit < j->size ? lo + unit : j->size;
        for (uint64_t i = lo; i < hi; ++i)
            t[j->index[i]] += j->value[i];
    }
    return NULL;
}
static void *expand_fold_worker(void *arg) {
    struct expand_job *j = (struct expand_job *) arg;
    const uint32_t unit = 16384;
    for (uint64_t lo = (uint64_t) j->worker * unit; lo < j->target_size; lo += (uint64_t) j->workers * unit) {
        uint64_t hi = lo + unit < j->target_size ? lo + unit : j->target_size;
        for (uint32_t w = 1; w < j->workers; ++w)
            for (uint64_t k = lo; k < hi; ++k)
                j->copies[k] += j->copies[k + (size_t) w * j->target_size];
    }
    return NULL;
}
/* scratch: workers * target_size floats owned by the caller (allocated once, like the expanded
 * variable that stays alive between kernel launches); indices must be < target_size. */
EXPORT int oracle_scatter_add_expand_f32(float *target, uint32_t target_size, const float *value,
                                         const uint32_t *index, uint32_t size, uint32_t workers,
                                         float *scratch) {
    if (workers == 0 || workers > 1024) return -1;
    pthread_t th[1024]; struct expand_job jobs[1024];
    memcpy(scratch, target, sizeof(float) * (size_t) target_size);
    memset(scratch + target_size, 0, sizeof(float) * (size_t) target_size * (workers - 1));
    for (int phase = 0; phase < 2; ++phase) {
        for (uint32_t w = 0; w < workers; ++w) {
            jobs[w] = (struct expand_job) { scratch, value, index, target_size, size, w, workers };
            pthread_create(&th[w], NULL, phase == 0 ? expand_scatter_worker : expand_fold_worker, &jobs[w]);
        }
        for (uint32_t w = 0; w < workers; ++w) pthread_join(th[w], NULL);
    }
    memcpy(target, scratch, sizeof(float) * (size_t) target_size);
    return 0;
}

/* memset_async: src/llvm_ts.cpp:215-263 / cuda_ts.cpp:129-183 (isize in {1,2,4,8}) */
EXPORT int oracle_memset(void *ptr, uint32_t size, uint32_t isize, const void *src) {
    if (isize != 1 && isize != 2 && isize != 4 && isize != 8) return -1;
    for (uint32_t i = 0; i < size; ++i) memcpy((uint8_t *) ptr + (size_t) i * isize, src, isize);
    return 0;
}

/* AggregationEntry: include/drjit-core/jit.h:2435-2443; kernels: resources/misc.cuh:31-61
 * and src/llvm_ts.cpp:976-992 (both agree; the header comment has the sign reversed):
 * size > 0: store the literal held in the `src` field itself (low `size` bytes);
 * size < 0: copy |size| bytes from the address `src`. Other sizes are ignored. */
struct AggregationEntry { int16_t size; uint16_t resource_kind; uint32_t offset; const void *src; };
EXPORT void oracle_aggregate(void *dst_, const struct AggregationEntry *agg, uint32_t n) {
    uint8_t *dst = (uint8_t *) dst_;
    for (uint32_t i = 0; i < n; ++i) {
        const struct AggregationEntry e = agg[i];
        int s = e.size;
        if (s == 1 || s == 2 || s == 4 || s == 8) {
            uint64_t lit = (uint64_t) (uintptr_t) e.src;
            memcpy(dst + e.offset, &lit, (size_t) s);      /* little endian */
        } else if (s == -1 || s == -2 || s == -4 || s == -8) {
            memcpy(dst + e.offset, e.src, (size_t) -s);
        }
    }
}
