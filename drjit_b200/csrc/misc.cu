/*
 * misc.cu -- fills, single-element stores, vcall parameter-block assembly, input generators.
 *
 * Replaces CUDAThreadState::memset_async (ext/drjit-core/src/cuda_ts.cpp:129-183), ::poke (:988-1006),
 * ::aggregate (:1008-1026) and the kernels fill_64 / poke_u* / aggregate (resources/misc.cuh:12-61).
 */
#include "common.cuh"
#include "runtime.h"

#include <cstring>

namespace djb {

/// Pattern fill with 128-bit stores: `pattern` holds the element replicated to 16 bytes
__global__ void __launch_bounds__(256)
fill_kernel(uint8_t *ptr, uint64_t bytes, uint4 pattern, uint32_t isize) {
    // head up to 16-byte alignment, vector body, tail -- all expressed in elements of isize
    const uintptr_t addr = (uintptr_t) ptr;
    uint64_t head = (16 - (addr & 15)) & 15;
    if (head > bytes) head = bytes;
    const uint64_t nvec = (bytes - head) / 16, tail_start = head + nvec * 16;
    const uint64_t gtid = (uint64_t) blockIdx.x * 256 + threadIdx.x, gstride = (uint64_t) gridDim.x * 256;
    const uint8_t *pat = reinterpret_cast<const uint8_t *>(&pattern);

    // `head` is a multiple of isize because ptr is isize-aligned
    for (uint64_t i = gtid; i < head; i += gstride) ptr[i] = pat[i % isize];
    for (uint64_t i = tail_start + gtid; i < bytes; i += gstride) ptr[i] = pat[(i - tail_start) % isize];
    // rotate the pattern so that it lines up with the (isize-aligned, 16-byte-misaligned) body
    uint4 *body = reinterpret_cast<uint4 *>(ptr + head);
    for (uint64_t i = gtid; i < nvec; i += gstride) body[i] = pattern;
}

void memset_async(cudaStream_t stream, void *ptr, uint32_t size_, uint32_t isize, const void *src) {
    if (isize != 1 && isize != 2 && isize != 4 && isize != 8) // cuda_ts.cpp:132-133
        raise(DRJIT_B200_EINVAL, "jit_memset_async(): invalid element size (must be 1, 2, 4, or 8)!");
    if (size_ == 0)
        return;
    uint64_t bytes = (uint64_t) size_ * isize;

    // Try to convert into ordinary memset if possible (cuda_ts.cpp:143-148)
    const uint8_t *s = (const uint8_t *) src;
    bool uniform = true;
    for (uint32_t i = 1; i < isize; ++i)
        uniform &= s[i] == s[0];
    if (uniform) {
        DJB_CUDA_CHECK(cudaMemsetAsync(ptr, s[0], bytes, stream));
        return;
    }
    if (((uintptr_t) ptr % isize) != 0)
        raise(DRJIT_B200_EINVAL, "jit_memset_async(): unaligned destination pointer!");

    uint8_t pat[16];
    for (uint32_t i = 0; i < 16; ++i)
        pat[i] = s[i % isize];
    uint4 pattern;
    memcpy(&pattern, pat, 16);
    const DeviceProps &dev = device_props();
    const uint32_t grid = (uint32_t) std::max<uint64_t>(1, std::min<uint64_t>(ceil_div64(bytes, 256 * 64), dev.sm_count * 8));
    fill_kernel<<<grid, 256, 0, stream>>>((uint8_t *) ptr, bytes, pattern, isize);
    DJB_POST_LAUNCH();
}

__global__ void poke_kernel(void *dst, uint64_t value, uint32_t size) {
    switch (size) {
        case 1: *(uint8_t *) dst = (uint8_t) value; break;
        case 2: *(uint16_t *) dst = (uint16_t) value; break;
        case 4: *(uint32_t *) dst = (uint32_t) value; break;
        default: *(uint64_t *) dst = value; break;
    }
}

void poke(cudaStream_t stream, void *dst, const void *src, uint32_t size) {
    if (size != 1 && size != 2 && size != 4 && size != 8) // cuda_ts.cpp:997-999
        raise(DRJIT_B200_EINVAL, "jit_poke(): only size=1, 2, 4 or 8 are supported!");
    uint64_t value = 0;
    memcpy(&value, src, size);
    poke_kernel<<<1, 1, 0, stream>>>(dst, value, size);
    DJB_POST_LAUNCH();
}

/// resources/misc.cuh:41-61 -- positive size: literal stored in `src`; negative: dereference
__global__ void __launch_bounds__(128)
aggregate_kernel(uint8_t *out, const drjit_b200_aggregation_entry *in, uint32_t size) {
    const uint32_t idx = blockIdx.x * 128 + threadIdx.x;
    if (idx >= size)
        return;
    // one 16-byte load per record
    const uint4 raw = *reinterpret_cast<const uint4 *>(in + idx);
    const int16_t rsize = (int16_t) (raw.x & 0xffffu);
    const uint32_t offset = raw.y;
    const uint64_t srcbits = ((uint64_t) raw.w << 32) | raw.z;
    void *dst = out + offset;
    const void *src = (const void *) (uintptr_t) srcbits;
    switch (rsize) {
        case  1: *(uint8_t *)  dst = (uint8_t)  srcbits; break;
        case  2: *(uint16_t *) dst = (uint16_t) srcbits; break;
        case  4: *(uint32_t *) dst = (uint32_t) srcbits; break;
        case  8: *(uint64_t *) dst = (uint64_t) srcbits; break;
        case -1: *(uint8_t *)  dst = *(const uint8_t *)  src; break;
        case -2: *(uint16_t *) dst = *(const uint16_t *) src; break;
        case -4: *(uint32_t *) dst = *(const uint32_t *) src; break;
        case -8: *(uint64_t *) dst = *(const uint64_t *) src; break;
    }
}

void aggregate(cudaStream_t stream, void *dst, const drjit_b200_aggregation_entry *agg, uint32_t size) {
    static_assert(sizeof(drjit_b200_aggregation_entry) == 16, "AggregationEntry layout (jit.h:2435-2443)");
    if (size == 0)
        return;
    aggregate_kernel<<<ceil_div(size, 128), 128, 0, stream>>>((uint8_t *) dst, agg, size);
    DJB_POST_LAUNCH();
}

// ---------------------------------------------------------------------------
//  Synthetic benchmark inputs (generator: ext/drjit-core/tests/reductions.cpp:5-13)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fill_fmix32_kernel(int kind, void *out, uint64_t start, uint64_t n, uint32_t xor_, uint32_t and_) {
    const uint64_t gstride = (uint64_t) gridDim.x * 256;
    for (uint64_t i = (uint64_t) blockIdx.x * 256 + threadIdx.x; i < n; i += gstride) {
        const uint32_t h = fmix32((uint32_t) (start + i) ^ xor_);
        if (kind == 0) ((uint32_t *) out)[i] = h & and_;
        else if (kind == 1) ((float *) out)[i] = (float) (h >> 8) * (1.0f / 16777216.0f);
        else ((uint8_t *) out)[i] = (h & 0xffu) < and_ ? 1 : 0;
    }
}

void fill_fmix32(cudaStream_t stream, int kind, void *out, uint64_t start, uint64_t n, uint32_t xor_,
                 uint32_t and_) {
    if (kind < 0 || kind > 2)
        raise(DRJIT_B200_EINVAL, "drjit_b200_fill_fmix32(): invalid kind!");
    if (n == 0)
        return;
    const DeviceProps &dev = device_props();
    const uint32_t grid = (uint32_t) std::max<uint64_t>(1, std::min<uint64_t>(ceil_div64(n, 256 * 8), dev.sm_count * 16));
    fill_fmix32_kernel<<<grid, 256, 0, stream>>>(kind, out, start, n, xor_, and_);
    DJB_POST_LAUNCH();
}

} // namespace djb
