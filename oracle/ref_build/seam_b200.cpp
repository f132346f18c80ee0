/*
 * seam_b200.cpp -- the binding INTEGRATION.md describes, as compiled code: the eight primitive
 * methods of the reference's CUDAThreadState (ext/drjit-core/src/cuda_ts.cpp:129-183, :195-352,
 * :354-398, :530-681, :683-763, :788-975, :988-1006, :1008-1026) implemented on top of
 * libdrjit_b200.so through include/drjit_b200_thread_state.h.
 *
 * oracle/ref_build/Makefile (`make b200`) links this file into oracle/_ref_b200/libdrjit-core.so next
 * to the reference's unmodified cuda_ts.o whose eight primitive symbols were weakened (objcopy), so
 * these definitions win at link time; the reference's own tests/reductions.cpp and tests/vcall.cpp
 * and tests/test_insitu_gpu.py then run through jit_block_reduce / jit_compress /
 * jit_block_mkperm(JitBackend::CUDA) of that library on the B200 box.
 *
 * JitFlag::KernelHistory / LaunchBlocking (cuda_ts.cpp:19-46): kept alive through the library's
 * launch hook -- one KernelHistoryEntry per primitive call with the reference's KernelType,
 * bracketed by two CUDA events exactly like submit_gpu() does per launch.
 */
#include "cuda_ts.h"
#include "cuda.h"
#include "util.h"
#include "var.h"
#include "log.h"

#define DRJIT_B200_USE_JIT_H
#include <drjit_b200_thread_state.h>

namespace {

/// KernelHistory / LaunchBlocking at the seam; called on the thread that issued the primitive,
/// which holds state.lock (the reference calls submit_gpu() under the same lock)
void launch_hook(void * /*user*/, int phase, int kernel_type, uint32_t size, void *stream_,
                 uint32_t launches, void **cookie) {
    CUstream stream = (CUstream) stream_;
    uint32_t flags = jit_flags();
    if (phase == 0) {
        *cookie = nullptr;
        if (unlikely(flags & (uint32_t) JitFlag::KernelHistory)) {
            KernelHistoryEntry *e = new KernelHistoryEntry();
            cuda_check(cuEventCreate((CUevent *) &e->event_start, CU_EVENT_DEFAULT));
            cuda_check(cuEventCreate((CUevent *) &e->event_end, CU_EVENT_DEFAULT));
            cuda_check(cuEventRecord((CUevent) e->event_start, stream));
            *cookie = e;
        }
        return;
    }
    if (unlikely(flags & (uint32_t) JitFlag::LaunchBlocking) && launches)
        cuda_check(cuStreamSynchronize(stream));
    KernelHistoryEntry *e = (KernelHistoryEntry *) *cookie;
    if (!e)
        return;
    if (launches) {
        ThreadState *ts = thread_state_cuda;
        e->backend = JitBackend::CUDA;
        // primitives the reference runs inside JIT kernels (scatter-reduce, >= 256) are tagged JIT
        e->type = kernel_type < 256 ? (KernelType) kernel_type : KernelType::JIT;
        e->recording_mode = ts ? ts->recording_mode : (KernelRecordingMode) 0;
        e->size = size;
        e->input_count = 1;
        e->output_count = 1;
        cuda_check(cuEventRecord((CUevent) e->event_end, stream));
        state.kernel_history.append(*e);
    } else {
        cuda_check(cuEventDestroy((CUevent) e->event_start));
        cuda_check(cuEventDestroy((CUevent) e->event_end));
    }
    delete e;
    *cookie = nullptr;
}

struct InstallHook {
    InstallHook() { drjit_b200_set_launch_hook(launch_hook, nullptr); }
} install_hook;

drjit_b200::ThreadState seam(CUDAThreadState *ts) { return drjit_b200::ThreadState((void *) ts->stream); }

} // namespace

void CUDAThreadState::memset_async(void *ptr, uint32_t size, uint32_t isize, const void *src) {
    scoped_set_context guard(context);
    seam(this).memset_async(ptr, size, isize, src);
}

void CUDAThreadState::block_reduce(VarType vt, ReduceOp op, uint32_t size, uint32_t block_size,
                                   const void *in, void *out) {
    jitc_log(Debug, "jit_block_reduce(" DRJIT_PTR " -> " DRJIT_PTR ", type=%s, op=%s, size=%u, block_size=%u) [b200]",
             (uintptr_t) in, (uintptr_t) out, type_name[(int) vt], red_name[(int) op], size, block_size);
    scoped_set_context guard(context);
    seam(this).block_reduce(vt, op, size, block_size, in, out);
}

void CUDAThreadState::block_prefix_reduce(VarType vt, ReduceOp op, uint32_t size, uint32_t block_size,
                                          bool exclusive, bool reverse, const void *in, void *out) {
    jitc_log(Debug, "jit_block_prefix_reduce(" DRJIT_PTR " -> " DRJIT_PTR ", type=%s, op=%s, size=%u, block_size=%u) [b200]",
             (uintptr_t) in, (uintptr_t) out, type_name[(int) vt], red_name[(int) op], size, block_size);
    scoped_set_context guard(context);
    seam(this).block_prefix_reduce(vt, op, size, block_size, exclusive, reverse, in, out);
}

void CUDAThreadState::reduce_dot(VarType vt, const void *ptr_1, const void *ptr_2, uint32_t size, void *out) {
    scoped_set_context guard(context);
    seam(this).reduce_dot(vt, ptr_1, ptr_2, size, out);
}

uint32_t CUDAThreadState::compress(const uint8_t *in, uint32_t size, uint32_t *out) {
    if (size == 0)
        return 0;
    scoped_set_context guard(context);
    // blocks on the stream like the reference (cuda_ts.cpp:759); state.lock stays held here --
    // a maintainer may wrap the call in unlock_guard once the launch hook takes the lock itself
    uint32_t count = seam(this).compress(in, size, out);
    jitc_log(Debug, "jit_compress(" DRJIT_PTR " -> " DRJIT_PTR ", size=%u) = %u [b200]",
             (uintptr_t) in, (uintptr_t) out, size, count);
    return count;
}

uint32_t CUDAThreadState::block_mkperm(const uint32_t *values, uint32_t size, uint32_t block_size,
                                       uint32_t bucket_count, uint32_t *perm, uint32_t *offsets) {
    if (size == 0)
        return 0;
    scoped_set_context guard(context);
    uint32_t unique = seam(this).block_mkperm(values, size, block_size, bucket_count, perm, offsets);
    jitc_log(Debug, "jit_block_mkperm(" DRJIT_PTR ", size=%u, block_size=%u, bucket_count=%u) = %u [b200]",
             (uintptr_t) values, size, block_size, bucket_count, unique);
    return unique;
}

void CUDAThreadState::poke(void *dst, const void *src, uint32_t size) {
    scoped_set_context guard(context);
    seam(this).poke(dst, src, size);
}

void CUDAThreadState::aggregate(void *dst, AggregationEntry *agg, uint32_t size) {
    scoped_set_context guard(context);
    seam(this).aggregate(dst, agg, size);
}
