/*
 * extern "C" shim over the reference's (C++-mangled) jit_* entry points.
 * TEST INFRASTRUCTURE ONLY -- lets Python (ctypes) drive the UNMODIFIED
 * reference libdrjit-core.so built by this directory's Makefile:
 *   - JitBackend::LLVM  : CPU primitives (llvm_ts.cpp:265-933) = oracle pin + CPU baseline
 *   - JitBackend::CUDA  : the incumbent compute_75-PTX kernels (GPU box only)
 * Signatures: ext/drjit-core/include/drjit-core/jit.h:92,141,301,446,472,2203,2239,2365,2387,2426.
 */
#include <drjit-core/jit.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <exception>

#define SHIM extern "C" __attribute__((visibility("default")))

static char last_error[1024];

template <typename F> static int guard(F &&f) {
    try { f(); last_error[0] = 0; return 0; }
    catch (const std::exception &e) {
        snprintf(last_error, sizeof(last_error), "%s", e.what());
        return -1;
    }
}

SHIM const char *ref_last_error() { return last_error; }

/// returns bit mask of initialised backends (bit 1 = CUDA, bit 2 = LLVM)
SHIM int ref_init(int want_cuda, int want_llvm) {
    uint32_t mask = 0;
    if (want_cuda) mask |= 1u << (uint32_t) JitBackend::CUDA;
    if (want_llvm) mask |= 1u << (uint32_t) JitBackend::LLVM;
    jit_init(mask);
    int rv = 0;
    if (want_cuda && jit_has_backend(JitBackend::CUDA)) rv |= 1 << 1;
    if (want_llvm && jit_has_backend(JitBackend::LLVM)) rv |= 1 << 2;
    return rv;
}
SHIM void ref_shutdown() { jit_shutdown(1); }
SHIM void ref_llvm_set_thread_count(uint32_t n) { jit_llvm_set_thread_count(n); }
SHIM void ref_sync() { jit_sync_thread(); }
SHIM void *ref_malloc(int backend, size_t size, int shared) {
    return jit_malloc((JitBackend) backend, size, shared);
}
SHIM void ref_free(void *p) { jit_free(p); }
SHIM void ref_memcpy(int backend, void *dst, const void *src, size_t size) {
    jit_memcpy((JitBackend) backend, dst, src, size);
}
SHIM void *ref_cuda_stream() { return jit_cuda_stream(); }

SHIM int ref_block_reduce(int backend, int vt, int op, uint32_t size,
                          uint32_t block_size, const void *in, void *out) {
    return guard([&] {
        jit_block_reduce((JitBackend) backend, (VarType) vt, (ReduceOp) op,
                         size, block_size, in, out);
    });
}

/// NB: effective positional order of the reference entry is (size, block_size),
/// see api.cpp:1331-1337 -> util.cpp:55-61 (the header names them the other way round).
SHIM int ref_block_prefix_reduce(int backend, int vt, int op, uint32_t size,
                                 uint32_t block_size, int exclusive, int reverse,
                                 const void *in, void *out) {
    return guard([&] {
        jit_block_prefix_reduce((JitBackend) backend, (VarType) vt, (ReduceOp) op,
                                size, block_size, exclusive, reverse, in, out);
    });
}

SHIM int64_t ref_compress(int backend, const uint8_t *in, uint32_t size, uint32_t *out) {
    int64_t rv = -1;
    guard([&] { rv = jit_compress((JitBackend) backend, in, size, out); });
    return rv;
}

SHIM int64_t ref_block_mkperm(int backend, const uint32_t *values, uint32_t size,
                              uint32_t block_size, uint32_t bucket_count,
                              uint32_t *perm, uint32_t *offsets) {
    int64_t rv = -1;
    guard([&] {
        rv = jit_block_mkperm((JitBackend) backend, values, size, block_size,
                              bucket_count, perm, offsets);
    });
    return rv;
}

/// dot product: the only public entry is variable-level (jit.h:2177), so map
/// the two buffers as variables, reduce, and read back the scalar.
SHIM int ref_reduce_dot(int backend, int vt, void *a, void *b, uint32_t size, void *out) {
    return guard([&] {
        uint32_t va = jit_var_mem_map((JitBackend) backend, (VarType) vt, a, size, 0),
                 vb = jit_var_mem_map((JitBackend) backend, (VarType) vt, b, size, 0),
                 vr = jit_var_reduce_dot(va, vb);
        jit_var_read(vr, 0, out);
        jit_var_dec_ref(va); jit_var_dec_ref(vb); jit_var_dec_ref(vr);
    });
}

/// CUDA-only incumbent for the histogram row: dr.scatter_reduce(Add) through the
/// reference tracer (jit.h:1102; JIT-emits PTX, cuda_scatter.cpp:246-354).
/// `target`, `value`, `index` are device pointers; evaluation is forced, not synced.
SHIM int ref_scatter_reduce(int backend, int vt, int op, int mode, void *target, uint32_t target_size,
                            void *value, void *index, uint32_t size) {
    return guard([&] {
        JitBackend be = (JitBackend) backend;
        uint32_t vt_ = jit_var_mem_map(be, (VarType) vt, target, target_size, 0),
                 vv  = jit_var_mem_map(be, (VarType) vt, value, size, 0),
                 vi  = jit_var_mem_map(be, VarType::UInt32, index, size, 0),
                 vm  = jit_var_bool(be, true);
        uint32_t vr = jit_var_scatter(vt_, vv, vi, vm, (ReduceOp) op, (ReduceMode) mode);
        jit_var_eval(vr);
        void *p = nullptr;
        jit_var_data(vr, &p);
        if (p != target) // copy-on-write kicked in (should not: refcount is 1)
            jit_memcpy_async(be, target, p, (size_t) target_size * (vt == (int) VarType::Float64 || vt == (int) VarType::UInt64 || vt == (int) VarType::Int64 ? 8 : 4));
        jit_var_dec_ref(vr); jit_var_dec_ref(vv); jit_var_dec_ref(vi); jit_var_dec_ref(vm);
    });
}

/// JitFlag (jit.h:1680-1783), e.g. KernelHistory = 1 << 15, LaunchBlocking = 1 << 16
SHIM void ref_set_flag(uint32_t flag, int enable) { jit_set_flag((JitFlag) flag, enable); }

/// jit_kernel_history() (jit.h:2700-2736) flattened for ctypes: copies up to `max` entries
/// (backend, KernelType, size, execution time in ms) and frees the list. Returns the entry count.
SHIM uint32_t ref_kernel_history(uint32_t *backends, uint32_t *types, uint32_t *sizes, float *times,
                                 uint32_t max) {
    KernelHistoryEntry *data = jit_kernel_history();
    uint32_t n = 0;
    if (!data)
        return 0;
    for (KernelHistoryEntry *e = data; (uint32_t) e->backend; ++e) {
        if (n < max) {
            backends[n] = (uint32_t) e->backend; types[n] = (uint32_t) e->type;
            sizes[n] = e->size; times[n] = e->execution_time;
        }
        ++n;
        free(e->ir);
    }
    free(data);
    return n;
}
SHIM void ref_kernel_history_clear() { jit_kernel_history_clear(); }

/// dr.scatter_reduce through the reference tracer with an explicit mask array (u8 device array, one
/// entry per element, may be NULL) -- otherwise like ref_scatter_reduce.
SHIM int ref_scatter_reduce_masked(int backend, int vt, int op, int mode, void *target, uint32_t target_size,
                                   void *value, void *index, void *mask, uint32_t size) {
    return guard([&] {
        JitBackend be = (JitBackend) backend;
        uint32_t vt_ = jit_var_mem_map(be, (VarType) vt, target, target_size, 0),
                 vv  = jit_var_mem_map(be, (VarType) vt, value, size, 0),
                 vi  = jit_var_mem_map(be, VarType::UInt32, index, size, 0),
                 vm  = mask ? jit_var_mem_map(be, VarType::Bool, mask, size, 0) : jit_var_bool(be, true);
        uint32_t vr = jit_var_scatter(vt_, vv, vi, vm, (ReduceOp) op, (ReduceMode) mode);
        jit_var_eval(vr);
        void *p = nullptr;
        jit_var_data(vr, &p);
        if (p != target)
            jit_memcpy_async(be, target, p, (size_t) target_size * (vt == (int) VarType::Float64 || vt == (int) VarType::UInt64 || vt == (int) VarType::Int64 ? 8 : 4));
        jit_var_dec_ref(vr); jit_var_dec_ref(vv); jit_var_dec_ref(vi); jit_var_dec_ref(vm);
    });
}

/// dr.scatter_reduce of an n-component packet through the reference tracer (jit_var_scatter_packet,
/// jit.h:1117-1120; PTX from cuda_packet.cpp:168-327). `values` = host array of n device pointers,
/// `target` = target_packets * n elements; mask (u8 device array) may be NULL.
SHIM int ref_scatter_packet(int backend, int vt, int op, int mode, void *target, uint32_t target_packets,
                            void **values, uint32_t n, void *index, void *mask, uint32_t size) {
    return guard([&] {
        JitBackend be = (JitBackend) backend;
        const size_t tsize = vt == (int) VarType::Float64 || vt == (int) VarType::UInt64 || vt == (int) VarType::Int64 ? 8
                           : vt == (int) VarType::Float16 ? 2 : 4;
        uint32_t vt_ = jit_var_mem_map(be, (VarType) vt, target, (size_t) target_packets * n, 0),
                 vi  = jit_var_mem_map(be, VarType::UInt32, index, size, 0),
                 vm  = mask ? jit_var_mem_map(be, VarType::Bool, mask, size, 0) : jit_var_bool(be, true);
        uint32_t vv[16];
        for (uint32_t k = 0; k < n; ++k)
            vv[k] = jit_var_mem_map(be, (VarType) vt, values[k], size, 0);
        uint32_t vr = jit_var_scatter_packet(n, vt_, vv, vi, vm, (ReduceOp) op, (ReduceMode) mode);
        jit_var_eval(vr);
        void *p = nullptr;
        jit_var_data(vr, &p);
        if (p != target)
            jit_memcpy_async(be, target, p, (size_t) target_packets * n * tsize);
        jit_var_dec_ref(vr); jit_var_dec_ref(vi); jit_var_dec_ref(vm);
        for (uint32_t k = 0; k < n; ++k) jit_var_dec_ref(vv[k]);
    });
}

/// dr.scatter_inc through the reference tracer (jit_var_scatter_inc, jit.h:1141-1143; PTX from
/// cuda_scatter.cpp:356-393): out[i] = target[index[i]]++. All device pointers; mask may be NULL.
SHIM int ref_scatter_inc(int backend, void *target, uint32_t target_size, void *index, void *mask,
                         uint32_t size, void *out) {
    return guard([&] {
        JitBackend be = (JitBackend) backend;
        uint32_t vt_ = jit_var_mem_map(be, VarType::UInt32, target, target_size, 0),
                 vi  = jit_var_mem_map(be, VarType::UInt32, index, size, 0),
                 vm  = mask ? jit_var_mem_map(be, VarType::Bool, mask, size, 0) : jit_var_bool(be, true);
        uint32_t vr = jit_var_scatter_inc(&vt_, vi, vm);
        jit_var_eval(vr);
        void *p = nullptr;
        jit_var_data(vr, &p);
        jit_memcpy_async(be, out, p, (size_t) size * 4);
        void *pt = nullptr;
        jit_var_data(vt_, &pt);
        if (pt != target)
            jit_memcpy_async(be, target, pt, (size_t) target_size * 4);
        jit_sync_thread();
        jit_var_dec_ref(vr); jit_var_dec_ref(vt_); jit_var_dec_ref(vi); jit_var_dec_ref(vm);
    });
}

/// Number of kernel-history entries whose IR (the PTX the JIT generated) contains `needle`;
/// consumes the history like ref_kernel_history.
SHIM uint32_t ref_kernel_history_ir_count(const char *needle) {
    KernelHistoryEntry *data = jit_kernel_history();
    uint32_t n = 0;
    if (!data)
        return 0;
    for (KernelHistoryEntry *e = data; (uint32_t) e->backend; ++e) {
        if (e->ir && strstr(e->ir, needle)) ++n;
        free(e->ir);
    }
    free(data);
    return n;
}
