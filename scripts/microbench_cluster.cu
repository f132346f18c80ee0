// Micro-benchmark for the next step of the mkperm scatter pass (lever (b)):
// what does the copy-out cost when a cluster of C CTAs merges the runs of C consecutive tiles
// through distributed shared memory, so that a bucket's entries leave the SM as runs of C*R words
// instead of R words?  (profiles/r2q_mkperm_phases.txt: the global stores of the copy-out are the
// largest single term of the pass, 104 of 338 us at R = 12.)
//
//   build/microbench_cluster            (make -C scripts)
//
// Every CTA owns one tile of T = 1024 * KPT staged entries in shared memory, ordered by bucket
// (B buckets, R = T / B entries per bucket and tile -- uniform keys). Tiles are walked in the same
// order by all CTAs, as in the shipped kernel. Output layout = the real one: bucket-major, inside a
// bucket tile-major, every bucket region misaligned by a few words.
//   C = 1 : thread j copies slot j of its own tile             (the shipped copy-out, minus `delta`)
//   C > 1 : CTA r of a cluster copies the buckets [r*B/C, (r+1)*B/C) of all C tiles: merged slot m
//           -> (bucket, tile q, j); the entry is read from CTA q's shared memory.
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr uint32_t kThreads = 1024;

struct Params {
    uint32_t *out;
    uint32_t tiles;         // total number of tiles (multiple of the cluster size)
    uint32_t stride;        // words per bucket region (tiles * R + pad)
};

template <uint32_t KPT, uint32_t B, uint32_t C>
__global__ void __launch_bounds__(kThreads, 1) copyout(const Params p) {
    constexpr uint32_t T = kThreads * KPT, R = T / B;
    static_assert(T % B == 0 && B % C == 0, "uniform runs");
    extern __shared__ __align__(16) uint32_t sorted[];                  // [T]
    const uint32_t tid = threadIdx.x;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = C > 1 ? cluster.block_rank() : 0;
    const uint32_t first = (blockIdx.x / C) * C;                        // first tile of the cluster in round 0

    for (uint32_t t0 = first; t0 < p.tiles; t0 += gridDim.x) {
        // stand-in for the ranking phase: slot j of this CTA's tile holds a local index
        for (uint32_t j = tid; j < T; j += kThreads) sorted[j] = (j * 2654435761u) >> 16;
        if (C > 1) cluster.sync(); else __syncthreads();

        if (C == 1) {
            const uint32_t idx0 = t0 * T;
            #pragma unroll 4
            for (uint32_t j = tid; j < T; j += kThreads) {
                const uint32_t b = j / R, w = j - b * R;
                p.out[(size_t) b * p.stride + (b * 3u & 7u) + t0 * R + w] = idx0 + sorted[j];
            }
        } else {
            constexpr uint32_t BPC = B / C;                             // buckets per CTA
            const uint32_t b_lo = rank * BPC;
            const uint32_t *remote[C];
            #pragma unroll
            for (uint32_t q = 0; q < C; ++q) remote[q] = cluster.map_shared_rank(sorted, q);
            #pragma unroll 4
            for (uint32_t m = tid; m < T; m += kThreads) {
                const uint32_t b = b_lo + m / (C * R), rest = m % (C * R), q = rest / R, w = rest - q * R;
                const uint32_t *src = remote[0];
                #pragma unroll
                for (uint32_t i = 1; i < C; ++i) if (q == i) src = remote[i];
                p.out[(size_t) b * p.stride + (b * 3u & 7u) + t0 * R + rest] = (t0 + q) * T + src[b * R + w];
            }
        }
        if (C > 1) cluster.sync(); else __syncthreads();
    }
}

template <uint32_t KPT, uint32_t B, uint32_t C>
static void run(int sms, uint32_t *out, size_t out_words, uint32_t n_log2) {
    constexpr uint32_t T = kThreads * KPT, R = T / B;
    Params p{};
    p.out = out;
    p.tiles = (uint32_t) (((1ull << n_log2) / T) / C * C);
    p.stride = p.tiles * R + 8;
    if ((size_t) B * p.stride > out_words) { printf("buffer too small\n"); return; }
    const uint32_t smem = T * 4;
    CK(cudaFuncSetAttribute(copyout<KPT, B, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((uint32_t) (sms / C * C)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = C; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int clusters = 0;
    CK(cudaOccupancyMaxActiveClusters(&clusters, copyout<KPT, B, C>, &cfg));
    if (C > 1 && (uint32_t) clusters * C < cfg.gridDim.x) cfg.gridDim = dim3(clusters * C);   // co-resident clusters only
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int i = 0; i < 6; ++i) {
        CK(cudaEventRecord(a));
        CK(cudaLaunchKernelEx(&cfg, copyout<KPT, B, C>, p));
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (i && ms < best) best = ms;
    }
    const double bytes = (double) p.tiles * T * 4;
    printf("tile %2u Ki entries, %4u buckets, cluster %u (grid %3u): runs of %3u words  %8.3f ms  %7.1f GB/s of permutation entries\n",
           T / 1024, B, C, cfg.gridDim.x, R * C, best, bytes / best / 1e6);
}

// Rate of scattered 4-byte shared-memory stores when a fraction of them goes to the other CTAs of
// the cluster (option (ii) of DESIGN.md section 8.1: entries ranked locally, stored straight into the
// owner CTA's staging buffer). MODE 0: plain stores, 1: returning atomicAdd on the (remote) word.
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h;
}

template <uint32_t C, int MODE>
__global__ void __launch_bounds__(kThreads, 1) remote_scatter(uint32_t *sink, uint32_t words_mask, uint32_t iters) {
    extern __shared__ __align__(16) uint32_t buf[];
    cg::cluster_group cluster = cg::this_cluster();
    for (uint32_t j = threadIdx.x; j <= words_mask; j += kThreads) buf[j] = 0;
    uint32_t *peer[C];
    #pragma unroll
    for (uint32_t q = 0; q < C; ++q) peer[q] = C > 1 ? cluster.map_shared_rank(buf, q) : buf;
    if (C > 1) cluster.sync(); else __syncthreads();
    uint32_t acc = 0;
    const uint32_t seed = (blockIdx.x * kThreads + threadIdx.x) * iters;
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t h = fmix32(seed + it);
        uint32_t *dst = peer[0];
        #pragma unroll
        for (uint32_t q = 1; q < C; ++q) if ((h >> 28) % C == q) dst = peer[q];     // uniform choice of the owner
        if (MODE == 0) dst[h & words_mask] = it;
        else acc += atomicAdd(dst + (h & words_mask), 1u);
    }
    if (C > 1) cluster.sync(); else __syncthreads();
    if (acc == 0x12345678u) sink[0] = acc + buf[threadIdx.x];
}

template <uint32_t C, int MODE>
static void run_remote(int sms, uint32_t *sink, int clk_khz) {
    const uint32_t words = 32768, smem = words * 4, iters = 4096;
    CK(cudaFuncSetAttribute(remote_scatter<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((uint32_t) (sms / C * C)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = C; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int clusters = 0;
    CK(cudaOccupancyMaxActiveClusters(&clusters, remote_scatter<C, MODE>, &cfg));
    if (C > 1 && (uint32_t) clusters * C < cfg.gridDim.x) cfg.gridDim = dim3(clusters * C);
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int i = 0; i < 4; ++i) {
        CK(cudaEventRecord(a));
        CK(cudaLaunchKernelEx(&cfg, remote_scatter<C, MODE>, sink, words - 1, iters));
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (i && ms < best) best = ms;
    }
    const double ops = (double) cfg.gridDim.x * kThreads * iters;
    printf("scattered 4-byte %s into 128 KB, cluster %u (%u%% remote, grid %3u): %8.3f ms  %6.2f per clk and SM\n",
           MODE ? "shared atomicAdd (ret)" : "shared stores         ", C, C > 1 ? 100 * (C - 1) / C : 0, cfg.gridDim.x, best,
           ops / (best * 1e-3) / cfg.gridDim.x / (clk_khz * 1e3));
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t out_words = (size_t) 80 << 20;                         // 2^26 entries + padding
    uint32_t *out; CK(cudaMalloc(&out, out_words * 4)); CK(cudaMemset(out, 0, out_words * 4));
    const uint32_t n_log2 = 26;
    run<32, 4096, 1>(sms, out, out_words, n_log2);
    run<32, 4096, 2>(sms, out, out_words, n_log2);
    run<32, 4096, 4>(sms, out, out_words, n_log2);
    run<48, 4096, 1>(sms, out, out_words, n_log2);
    run<48, 4096, 2>(sms, out, out_words, n_log2);
    run<48, 4096, 4>(sms, out, out_words, n_log2);
    run<32, 256, 1>(sms, out, out_words, n_log2);                       // long runs: the streaming limit of this loop
    CK(cudaFree(out));

    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    uint32_t *sink; CK(cudaMalloc(&sink, 64));
    run_remote<1, 0>(sms, sink, clk); run_remote<2, 0>(sms, sink, clk); run_remote<4, 0>(sms, sink, clk);
    run_remote<1, 1>(sms, sink, clk); run_remote<2, 1>(sms, sink, clk);
    return 0;
}
