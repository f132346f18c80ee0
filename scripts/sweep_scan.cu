// sweep_scan.cu -- geometry sweep of the single-pass scan kernel (developer tool).
//   make -C scripts && build/sweep_scan [log2_n]
// Instantiates drjit_b200/csrc/scan_kernel.cuh for u32/Add with several (rows, stages, CTAs/SM)
// choices, checks every result on the device and prints CUDA-event timings.
#include "../drjit_b200/csrc/scan_kernel.cuh"

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

using namespace djb;

__global__ void fill(uint32_t *x, uint64_t n) {
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
        x[i] = fmix32((uint32_t) i);
}

__global__ void check(const uint32_t *x, const uint32_t *out, uint64_t n, unsigned long long *errors) {
    unsigned long long bad = 0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        if (i == 0) bad += out[0] != 0;
        else bad += (out[i] - out[i - 1]) != x[i - 1];
    }
    if (bad) atomicAdd(errors, bad);
}

static int g_sms = 0;
static uint32_t *g_in, *g_out;
static uint8_t *g_scratch;
static unsigned long long *g_err;
static const char *g_filter = nullptr;
static int g_debug = 0;

template <uint32_t R, uint32_t STAGES, uint32_t MIN_CTAS>
void run(uint64_t n, const char *label) {
    constexpr uint32_t CTAS_PER_SM = 0;
    using Geom = ScanGeom<uint32_t, true, R>;
    if (g_filter && !strstr(label, g_filter)) return;
    auto kernel = prefix_reduce_kernel<uint32_t, OpAdd, false, true, R, STAGES, MIN_CTAS>;
    constexpr uint32_t smem = STAGES * Geom::TILE_BYTES;
    constexpr uint32_t threads = ScanRoles<false, STAGES>::THREADS;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    if (occ < 1) { printf("%-34s does not fit\n", label); return; }
    if (CTAS_PER_SM && (int) CTAS_PER_SM < occ) occ = CTAS_PER_SM;
    cudaFuncAttributes attr; CK(cudaFuncGetAttributes(&attr, kernel));

    PrefixParams p{};
    p.in = g_in; p.out = g_out; p.size = (uint32_t) n; p.block_size = (uint32_t) n;
    p.exclusive = 1; p.debug = (uint8_t) g_debug; p.tiles = (uint32_t) ((n + Geom::TILE - 1) / Geom::TILE);
    p.state = g_scratch;
    const size_t state_bytes = (size_t) p.tiles * 8;
    const uint32_t grid = std::min<uint32_t>(p.tiles, g_sms * occ);

    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    std::vector<float> ts;
    for (int rep = 0; rep < 8; ++rep) {
        CK(cudaMemsetAsync(g_scratch, 0, state_bytes));
        CK(cudaEventRecord(a));
        void *args[] = { (void *) &p };
        CK(cudaLaunchCooperativeKernel((const void *) kernel, dim3(grid), dim3(threads), args, smem, 0));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep >= 3) ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    CK(cudaMemset(g_err, 0, 8));
    check<<<g_sms * 8, 256>>>(g_in, g_out, n, g_err);
    unsigned long long err = 0; CK(cudaMemcpy(&err, g_err, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemset(g_out, 0xff, n * 4));
    const float med = ts[ts.size() / 2];
    printf("%-34s tile %3u KiB regs %3d occ %d grid %4u  median %7.3f ms  %7.1f GB/s  %s\n", label,
           Geom::TILE_BYTES / 1024, attr.numRegs, occ, grid, med, n * 8.0 / med / 1e6, err ? "WRONG" : "ok");
    fflush(stdout);
}

int main(int argc, char **argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 30;
    const uint64_t n = 1ull << lg;
    if (argc > 2 && argv[2][0]) g_filter = argv[2];
    if (argc > 3) g_debug = atoi(argv[3]);
    CK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaMalloc(&g_in, n * 4)); CK(cudaMalloc(&g_out, n * 4));
    CK(cudaMalloc(&g_scratch, 64 << 20)); CK(cudaMalloc(&g_err, 8));
    fill<<<g_sms * 8, 256>>>(g_in, n);
    CK(cudaDeviceSynchronize());
    printf("exclusive u32 prefix sum, n = 2^%d, %d SMs, debug=%d\n", lg, g_sms, g_debug);

    // STAGES < 2: decoupled look-back; STAGES >= 2: early aggregates + windowed carry
    run<8, 0, 3>(n, "lookback direct R=8 min3");
    run<8, 1, 3>(n, "lookback staged R=8 S=1 min3");
    run<4, 2, 4>(n, "window R=4 S=2 min4");
    run<4, 3, 4>(n, "window R=4 S=3 min4");
    run<4, 4, 3>(n, "window R=4 S=4 min3");
    run<8, 2, 3>(n, "window R=8 S=2 min3");
    run<8, 2, 2>(n, "window R=8 S=2 min2");
    run<8, 3, 2>(n, "window R=8 S=3 min2");
    run<16, 2, 1>(n, "window R=16 S=2 min1");
    run<16, 3, 1>(n, "window R=16 S=3 min1");
    return 0;
}
