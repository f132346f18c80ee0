// sweep_compress.cu -- geometry sweep of the single-pass compress kernel (developer tool).
//   make -C scripts && build/sweep_compress [log2_n] [threshold 0..256]
#include "../drjit_b200/csrc/compress_kernel.cuh"

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

using namespace djb;

__global__ void fill(uint8_t *m, uint64_t n, uint32_t thr) {
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
        m[i] = (fmix32((uint32_t) i) & 0xffu) < thr ? 1 : 0;
}

// out must be strictly increasing, every out[j] must select a set byte, count must match popcount
__global__ void check(const uint8_t *m, const uint32_t *out, uint64_t n, const uint32_t *count,
                      unsigned long long *errors, unsigned long long *ones, uint32_t base) {
    unsigned long long bad = 0, c = 0;
    const uint32_t cnt = *count;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        c += m[i] != 0;
        if (i < cnt) {
            bad += (out[i] - base) >= n || m[out[i] - base] == 0;
            if (i > 0) bad += out[i] <= out[i - 1];
        }
    }
    if (bad) atomicAdd(errors, bad);
    if (c) atomicAdd(ones, c);
}

static int g_sms = 0;
static uint8_t *g_in, *g_scratch;
static uint32_t *g_out, *g_count;
static unsigned long long *g_err;
static double g_density;

static const char *g_filter = nullptr;
static int g_debug = 0;

template <uint32_t ROWS, uint32_t STAGES, uint32_t MIN_CTAS, uint32_t COPY = kCopyLsu, bool BASE512 = true>
void run(uint64_t n, const char *label) {
    constexpr uint32_t TILE = kCompThreads * ROWS * kCompUnit;
    constexpr uint32_t CTAS_PER_SM = 0;
    if (g_filter && !strstr(label, g_filter)) return;
    auto kernel = compress_kernel<ROWS, STAGES, MIN_CTAS, COPY, BASE512>;
    constexpr uint32_t smem = compress_smem_bytes<ROWS, STAGES, COPY>();
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kCompThreads, smem));
    if (occ < 1) { printf("%-34s does not fit\n", label); return; }
    if (CTAS_PER_SM && (int) CTAS_PER_SM < occ) occ = CTAS_PER_SM;
    cudaFuncAttributes attr; CK(cudaFuncGetAttributes(&attr, kernel));

    CompressParams p{};
    p.in = g_in; p.out = g_out; p.size = (uint32_t) n; p.index_base = BASE512 ? 0 : 7; p.debug = (uint32_t) g_debug;
    p.tiles = (uint32_t) ((n + TILE - 1) / TILE);
    p.state = (uint64_t *) g_scratch; p.count_out = g_count;
    const size_t state_bytes = (size_t) p.tiles * 8;
    const uint32_t grid = std::min<uint32_t>(p.tiles, g_sms * occ);

    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    std::vector<float> ts;
    for (int rep = 0; rep < 8; ++rep) {
        CK(cudaMemsetAsync(g_scratch, 0, state_bytes));
        CK(cudaEventRecord(a));
        void *args[] = { (void *) &p };
        CK(cudaLaunchCooperativeKernel((const void *) kernel, dim3(grid), dim3(kCompThreads), args, smem, 0));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep >= 3) ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    CK(cudaMemset(g_err, 0, 16));
    if (!g_debug) check<<<g_sms * 8, 256>>>(g_in, g_out, n, g_count, g_err, g_err + 1, p.index_base);
    unsigned long long res[2]; CK(cudaMemcpy(res, g_err, 16, cudaMemcpyDeviceToHost));
    uint32_t cnt; CK(cudaMemcpy(&cnt, g_count, 4, cudaMemcpyDeviceToHost));
    const bool ok = res[0] == 0 && res[1] == cnt;
    CK(cudaMemset(g_out, 0xff, n * 4)); CK(cudaMemset(g_count, 0, 4));
    const float med = ts[ts.size() / 2];
    printf("%-34s tile %3u KiB regs %3d occ %d grid %4u  median %7.3f ms  %7.1f GB/s  %s\n", label,
           TILE / 1024, attr.numRegs, occ, grid, med, n * (1.0 + 4.0 * g_density) / med / 1e6, ok ? "ok" : "WRONG");
    fflush(stdout);
}

int main(int argc, char **argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 30;
    const uint32_t thr = argc > 2 ? atoi(argv[2]) : 128;
    const uint64_t n = 1ull << lg;
    g_density = thr / 256.0;
    if (argc > 3 && argv[3][0]) g_filter = argv[3];
    if (argc > 4) g_debug = atoi(argv[4]);
    CK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaMalloc(&g_in, n)); CK(cudaMalloc(&g_out, n * 4));
    CK(cudaMalloc(&g_scratch, 64 << 20)); CK(cudaMalloc(&g_err, 16)); CK(cudaMalloc(&g_count, 4));
    fill<<<g_sms * 8, 256>>>(g_in, n, thr);
    CK(cudaMemset(g_count, 0, 4));
    CK(cudaDeviceSynchronize());
    printf("compress, n = 2^%d, density %.3f, %d SMs\n", lg, g_density, g_sms);

    run<8, 0, 3>(n, "direct ROWS=8 min3");
    run<4, 2, 4>(n, "ROWS=4 S=2 min4");
    run<4, 4, 3>(n, "ROWS=4 S=4 min3");
    run<8, 1, 3>(n, "ROWS=8 S=1 min3");
    run<8, 1, 3, kCopyLsuPairs>(n, "pairs ROWS=8 S=1 min3");
    run<8, 1, 4, kCopyLsuPairs>(n, "pairs ROWS=8 S=1 min4");
    run<8, 2, 3, kCopyLsuPairs>(n, "pairs ROWS=8 S=2 min3");
    run<16, 1, 2, kCopyLsuPairs>(n, "pairs ROWS=16 S=1 min2");
    run<8, 1, 3, kCopyVec>(n, "vec ROWS=8 S=1 min3");
    run<8, 1, 3, kCopyVec, false>(n, "vec ROWS=8 S=1 min3 base+7");
    run<8, 1, 4, kCopyVec>(n, "vec ROWS=8 S=1 min4");
    run<8, 2, 3, kCopyVec>(n, "vec ROWS=8 S=2 min3");
    run<4, 2, 4, kCopyVec>(n, "vec ROWS=4 S=2 min4");
    run<16, 1, 2, kCopyVec>(n, "vec ROWS=16 S=1 min2");
    run<8, 1, 3, kCopyBulk>(n, "bulk ROWS=8 S=1 min3");
    run<8, 1, 3, kCopyBulk, false>(n, "bulk ROWS=8 S=1 min3 base+7");
    run<8, 1, 2, kCopyBulk>(n, "bulk ROWS=8 S=1 min2");
    run<8, 2, 2, kCopyBulk>(n, "bulk ROWS=8 S=2 min2");
    run<8, 0, 3, kCopyBulk>(n, "bulk direct ROWS=8 min3");
    run<4, 2, 4, kCopyBulk>(n, "bulk ROWS=4 S=2 min4");
    run<16, 1, 2, kCopyBulk>(n, "bulk ROWS=16 S=1 min2");
    run<8, 1, 4>(n, "ROWS=8 S=1 min4");
    run<8, 2, 3>(n, "ROWS=8 S=2 min3");
    run<8, 2, 2>(n, "ROWS=8 S=2 min2");
    run<8, 3, 2>(n, "ROWS=8 S=3 min2");
    run<16, 1, 2>(n, "ROWS=16 S=1 min2");
    run<16, 2, 1>(n, "ROWS=16 S=2 min1");
    return 0;
}
