// microbench_partition.cu -- optimistic bound of "partition by the top 5 bits of the bin index, then
// accumulate every partition in privatised shared-memory bins" for dr.scatter_reduce(Add) of 2^28 f32
// values into 2^20 bins, against the 1.38 ms of the direct RED kernel (scatter_reduce.cu).
// Developer tool:  make -C scripts  &&  build/microbench_partition
//
// Three numbers, all at 2^28 (index, value) pairs:
//   (A) pass-A floor: read index + value (8 B), write 16-bit local bin + value (6 B) in the SAME order,
//       i.e. the memory traffic of the partition pass without any of its ranking work;
//   (B) pass B: stream the 6-byte records of an (artificially) already partitioned array, 2^15 bins per
//       partition in shared memory, shared-memory atomics, bins flushed with one RED each;
//   (C) for reference, the plain RED loop on the same data.
// (A) + (B) is what the scheme costs if the multi-split itself were free; the repository's own stable
// multi-split (sort.cu: 0.38 ms per 2^26 keys, profiles/r4d_prims_sort.txt) shows what it is not.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h += 1; h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h;
}

constexpr uint32_t kBinsLog2 = 20, kPartBits = 5, kLocalBits = kBinsLog2 - kPartBits, kLocalBins = 1u << kLocalBits;

__global__ void fill(uint32_t *idx, float *val, uint64_t n) {
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        idx[i] = fmix32((uint32_t) i ^ 0x85ebca6bu) & ((1u << kBinsLog2) - 1);
        val[i] = (float) (fmix32((uint32_t) i) >> 8) * (1.0f / 16777216.0f);
    }
}
// records "already partitioned": partition p owns the contiguous range [p n/32, (p+1) n/32); its records
// carry random local bins (what a real partition pass would have produced, up to order)
__global__ void fill_partitioned(uint16_t *loc, float *val, uint64_t n) {
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        loc[i] = (uint16_t) (fmix32((uint32_t) i ^ 0x85ebca6bu) & (kLocalBins - 1));
        val[i] = (float) (fmix32((uint32_t) i) >> 8) * (1.0f / 16777216.0f);
    }
}

// (A) traffic of the partition pass only: 8 B in, 6 B out per pair, vectorised
__global__ void __launch_bounds__(256) pass_a_floor(const uint4 *idx, const uint4 *val, uint2 *loc, uint4 *out_val, uint64_t nvec) {
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint4 k = __ldcs(idx + i), v = __ldcs(val + i);
        const uint32_t m = kLocalBins - 1;
        __stcs(loc + i, make_uint2((k.x & m) | ((k.y & m) << 16), (k.z & m) | ((k.w & m) << 16)));
        __stcs(out_val + i, v);
    }
}

// (B) one CTA = one slice of one partition: 2^15 f32 bins in shared memory
__global__ void __launch_bounds__(1024, 1) pass_b(const uint16_t *loc, const float *val, float *bins, uint64_t n, uint32_t slices) {
    extern __shared__ float sbins[];
    const uint32_t part = blockIdx.x / slices, slice = blockIdx.x % slices;
    for (uint32_t i = threadIdx.x; i < kLocalBins; i += blockDim.x) sbins[i] = 0.f;
    __syncthreads();
    const uint64_t per_part = n >> kPartBits, per_slice = per_part / slices;
    const uint64_t lo = (uint64_t) part * per_part + (uint64_t) slice * per_slice, hi = lo + per_slice;
    const uint4 *l4 = reinterpret_cast<const uint4 *>(loc + lo);        // 8 records per 16 bytes
    const uint4 *v4 = reinterpret_cast<const uint4 *>(val + lo);
    const uint64_t nvec = (hi - lo) / 8;
    for (uint64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
        const uint4 l = __ldcs(l4 + i), a = __ldcs(v4 + 2 * i), b = __ldcs(v4 + 2 * i + 1);
        atomicAdd(sbins + (l.x & 0xffffu), __uint_as_float(a.x)); atomicAdd(sbins + (l.x >> 16), __uint_as_float(a.y));
        atomicAdd(sbins + (l.y & 0xffffu), __uint_as_float(a.z)); atomicAdd(sbins + (l.y >> 16), __uint_as_float(a.w));
        atomicAdd(sbins + (l.z & 0xffffu), __uint_as_float(b.x)); atomicAdd(sbins + (l.z >> 16), __uint_as_float(b.y));
        atomicAdd(sbins + (l.w & 0xffffu), __uint_as_float(b.z)); atomicAdd(sbins + (l.w >> 16), __uint_as_float(b.w));
    }
    __syncthreads();
    float *dst = bins + (size_t) part * kLocalBins;
    for (uint32_t i = threadIdx.x; i < kLocalBins; i += blockDim.x)
        atomicAdd(dst + i, sbins[i]);
}

// (C) the direct loop: one RED per pair
__global__ void __launch_bounds__(256) direct_red(const uint4 *idx, const uint4 *val, float *bins, uint64_t nvec) {
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint4 k = __ldcs(idx + i), v = __ldcs(val + i);
        atomicAdd(bins + k.x, __uint_as_float(v.x)); atomicAdd(bins + k.y, __uint_as_float(v.y));
        atomicAdd(bins + k.z, __uint_as_float(v.z)); atomicAdd(bins + k.w, __uint_as_float(v.w));
    }
}

template <typename F> static float time_ms(F &&launch) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    std::vector<float> ts;
    for (int rep = 0; rep < 8; ++rep) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep >= 3) ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}

int main(int argc, char **argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 28;
    const uint64_t n = 1ull << lg;
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint32_t *idx; float *val, *val2, *bins; uint16_t *loc;
    CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&val, n * 4)); CK(cudaMalloc(&val2, n * 4)); CK(cudaMalloc(&loc, n * 2));
    CK(cudaMalloc(&bins, 4u << kBinsLog2)); CK(cudaMemset(bins, 0, 4u << kBinsLog2));
    fill<<<sms * 8, 256>>>(idx, val, n);
    CK(cudaDeviceSynchronize());
    printf("scatter_add of 2^%d f32 into 2^%u bins, %d SMs\n", lg, kBinsLog2, sms);

    const float ta = time_ms([&] { pass_a_floor<<<sms * 8, 256>>>((const uint4 *) idx, (const uint4 *) val, (uint2 *) loc, (uint4 *) val2, n / 4); });
    printf("(A) partition-pass traffic only (8 B in, 6 B out)      %7.3f ms  %7.1f GB/s\n", ta, n * 14.0 / ta / 1e6);

    fill_partitioned<<<sms * 8, 256>>>(loc, val2, n);
    CK(cudaDeviceSynchronize());
    CK(cudaFuncSetAttribute(pass_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kLocalBins * 4)));
    for (uint32_t slices : { 4u, 8u, 16u }) {
        const float tb = time_ms([&] { pass_b<<<(1u << kPartBits) * slices, 1024, kLocalBins * 4>>>(loc, val2, bins, n, slices); });
        printf("(B) accumulate, %2u CTAs per partition (6 B in)          %7.3f ms  %7.1f GB/s   (A)+(B) = %.3f ms\n",
               slices, tb, n * 6.0 / tb / 1e6, ta + tb);
    }
    const float tc = time_ms([&] { direct_red<<<sms * 32, 256>>>((const uint4 *) idx, (const uint4 *) val, bins, n / 4); });
    printf("(C) direct red.global.add.f32 per pair (8 B in)         %7.3f ms  %7.1f GB/s\n", tc, n * 8.0 / tc / 1e6);
    return 0;
}
