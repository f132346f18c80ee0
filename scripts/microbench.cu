// microbench.cu -- hardware rates that bound the atomic-heavy primitives (mkperm, scatter-add).
// Developer tool, not part of the product:  make -C scripts  &&  build/microbench
//
// Keys are generated in registers (fmix32 of the element index), so every number below is the
// rate of the instruction under test, not of HBM.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h += 1; h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h;
}

// mode 0: smem atomicAdd without return, 1: with return, 2: match.any + popc (+1 smem add by leader),
// 3: plain smem store (upper bound of the LSU path), 4: smem load (gather)
template <int MODE>
__global__ void smem_rate(uint32_t *sink, uint32_t bins_mask, uint32_t iters) {
    extern __shared__ uint32_t bins[];
    for (uint32_t i = threadIdx.x; i <= bins_mask; i += blockDim.x) bins[i] = 0;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 7919u;
    for (uint32_t it = 0; it < iters; ++it) {
        uint32_t k[8];
        #pragma unroll
        for (int u = 0; u < 8; ++u) k[u] = fmix32(seed + it * 8 + u) & bins_mask;
        #pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) atomicAdd(&bins[k[u]], 1u);
            else if (MODE == 1) acc += atomicAdd(&bins[k[u]], 1u);
            else if (MODE == 2) {
                uint32_t peers = __match_any_sync(0xffffffffu, k[u]);
                uint32_t lt; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
                acc += __popc(peers & lt);
                if ((peers & lt) == 0) bins[k[u]] += __popc(peers);
            }
            else if (MODE == 3) bins[k[u]] = it;
            else if (MODE == 4) acc += bins[k[u]];
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i <= bins_mask; i += blockDim.x) acc += bins[i];
    if (acc == 0x12345678u) sink[0] = acc;
}

// L2 atomics: red.global.add.f32 (no return) or u32 atomicAdd with return on random bins
template <int MODE>
__global__ void l2_rate(float *bins, uint32_t *sink, uint32_t bins_mask, uint32_t iters) {
    uint32_t acc = 0;
    uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 7919u;
    for (uint32_t it = 0; it < iters; ++it) {
        #pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t k = fmix32(seed + it * 8 + u) & bins_mask;
            if (MODE == 0) atomicAdd(&bins[k], 1.0f);
            else if (MODE == 1) acc += atomicAdd((uint32_t *) &bins[k], 1u);
            else if (MODE == 2) ((uint32_t *) bins)[k] = it;          // random 4-byte stores
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// Stores of `run` consecutive 4-byte words at random run-aligned places of a large buffer:
// the DRAM cost of scattered perm writes as a function of the run length
__global__ void run_store(uint32_t *buf, uint64_t words_mask, uint32_t run_log2, uint32_t iters) {
    const uint32_t lane = threadIdx.x & 31u, run = 1u << run_log2;
    uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint32_t it = 0; it < iters; ++it) {
        // each group of `run` lanes writes one run
        uint32_t g = lane >> run_log2;
        uint64_t r = ((uint64_t) fmix32((warp_global * iters + it) * 32u + g) << 3) ^ fmix32(it + g * 77u);
        uint64_t base = (r << run_log2) & words_mask;
        buf[base + (lane & (run - 1))] = it;
    }
}

template <typename F> float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0, clk = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("SMs %d, max clock %d kHz\n", sms, clk);
    uint32_t *sink; CK(cudaMalloc(&sink, 64));

    const char *names[] = { "smem atomicAdd (no ret)", "smem atomicAdd (ret)", "match.any+popc+leader add", "smem store", "smem load" };
    for (uint32_t bins : { 256u, 4096u, 16384u }) {
        for (int threads : { 256, 1024 }) {
            for (int mode = 0; mode < 5; ++mode) {
                const uint32_t iters = 2048;
                const int ctas_per_sm = threads == 256 ? 4 : 1;
                const size_t smem = bins * 4;
                const int grid = sms * ctas_per_sm;
                auto launch = [&]() {
                    switch (mode) {
                        case 0: smem_rate<0><<<grid, threads, smem>>>(sink, bins - 1, iters); break;
                        case 1: smem_rate<1><<<grid, threads, smem>>>(sink, bins - 1, iters); break;
                        case 2: smem_rate<2><<<grid, threads, smem>>>(sink, bins - 1, iters); break;
                        case 3: smem_rate<3><<<grid, threads, smem>>>(sink, bins - 1, iters); break;
                        default: smem_rate<4><<<grid, threads, smem>>>(sink, bins - 1, iters); break;
                    }
                };
                float ms = time_ms(launch);
                double keys = (double) grid * threads * iters * 8;
                printf("bins %5u thr %4d x%d/SM  %-28s %8.3f ms  %7.2f Gkeys/s  %5.2f keys/clk/SM (at %.2f GHz)\n",
                       bins, threads, ctas_per_sm, names[mode], ms, keys / ms / 1e6, keys / ms / 1e6 / sms / (clk / 1e6), clk / 1e6);
            }
        }
    }

    for (uint32_t lg : { 12u, 20u, 24u }) {
        float *bins; CK(cudaMalloc(&bins, (size_t) 4 << lg)); CK(cudaMemset(bins, 0, (size_t) 4 << lg));
        const char *n2[] = { "red.global.add.f32", "atom.global.add.u32 (ret)", "st.global random 4B" };
        for (int mode = 0; mode < 3; ++mode) {
            const uint32_t iters = 256; const int threads = 256, grid = sms * 8;
            auto launch = [&]() {
                if (mode == 0) l2_rate<0><<<grid, threads>>>(bins, sink, (1u << lg) - 1, iters);
                else if (mode == 1) l2_rate<1><<<grid, threads>>>(bins, sink, (1u << lg) - 1, iters);
                else l2_rate<2><<<grid, threads>>>(bins, sink, (1u << lg) - 1, iters);
            };
            float ms = time_ms(launch);
            double keys = (double) grid * threads * iters * 8;
            printf("L2 bins 2^%u  %-28s %8.3f ms  %7.2f Gops/s\n", lg, n2[mode], ms, keys / ms / 1e6);
        }
        CK(cudaFree(bins));
    }

    {
        const uint64_t words = 1ull << 28;  // 1 GiB buffer (> L2)
        uint32_t *buf; CK(cudaMalloc(&buf, words * 4)); CK(cudaMemset(buf, 0, words * 4));
        for (uint32_t run_log2 = 0; run_log2 <= 5; ++run_log2) {
            const uint32_t iters = 1024; const int threads = 256, grid = sms * 8;
            float ms = time_ms([&]() { run_store<<<grid, threads>>>(buf, words - 1, run_log2, iters); });
            double bytes = (double) grid * threads * iters * 4;
            printf("random runs of %2u words into 1 GiB: %8.3f ms  %7.1f GB/s of useful stores\n", 1u << run_log2, ms, bytes / ms / 1e6);
        }
        CK(cudaFree(buf));
    }
    return 0;
}
