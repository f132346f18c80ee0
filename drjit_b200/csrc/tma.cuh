/*
 * tma.cuh -- bulk asynchronous copies (TMA, 1-D form) and mbarrier helpers for sm_100a.
 *
 * The streaming primitives stage their input tiles in shared memory with
 * `cp.async.bulk.shared.global` (SASS: UBLKCP) several tiles ahead of the consumer warps, so
 * that the bytes in flight per SM are decoupled from the register footprint and from the
 * serial phases of a tile (look-back, combine, store). Completion is tracked with one
 * transaction-counting mbarrier per stage.
 */
#pragma once

#include <stdint.h>

namespace djb {

__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}

/// Makes mbarrier initialisation (generic proxy) visible to the async proxy
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}

/// Plain arrival (count 1) by the calling thread; earlier shared-memory writes of the thread
/// are visible to whoever observes the phase completion (release semantics at CTA scope)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}

/// L2 eviction policy for data that is touched exactly once
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

/// global -> shared bulk copy; `bytes` multiple of 16, both addresses 16-byte aligned.
/// Completion is signalled on `bar` as `bytes` transaction units.
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                          uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
                 "[%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
                 : "memory");
}

/// Asks L2 to fetch `bytes` (multiple of 16) starting at the 16-byte aligned address `gmem`:
/// one instruction, no destination, no completion tracking
__device__ __forceinline__ void bulk_prefetch_l2(const void *gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(gmem), "r"(bytes) : "memory");
}

/// shared -> global bulk copy (SASS: UBLKCP.G.S); `bytes` multiple of 16, both addresses 16-byte
/// aligned. Completion is tracked per thread through bulk groups (commit / wait below); the
/// shared-memory source must have been written before a fence.proxy.async by its writers.
__device__ __forceinline__ void bulk_store(void *gmem_dst, uint32_t smem_src_addr, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gmem_dst), "r"(smem_src_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
/// Blocks until all but the N most recent bulk groups of this thread have finished READING their
/// shared-memory source (the source may then be overwritten; the global writes may still be in flight)
template <uint32_t N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}

/// 128-bit shared-memory load
__device__ __forceinline__ uint4 lds128(const void *p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_addr(p)));
    return r;
}

} // namespace djb
