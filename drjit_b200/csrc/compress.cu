/*
 * compress.cu -- mask compaction: indices of the non-zero bytes of a mask, ascending.
 *
 * Replaces CUDAThreadState::compress (ext/drjit-core/src/cuda_ts.cpp:683-763) and the kernels
 * compress_small / compress_large / compress_large_init (resources/compress.cuh:23-156).
 *
 * Design: compress_kernel.cuh. Differences to the reference that are visible at the seam:
 *  - the mask is never written (the reference zero-pads the caller's buffer up to the next
 *    multiple of 2048, cuda_ts.cpp:746-748); the ragged tail is bounds-checked instead;
 *  - the count goes to a device-mapped pinned word; one launch + one memset in total (the
 *    reference: init kernel + main kernel + 14 block-wide barriers per 2048 mask bytes and
 *    scattered 4-byte index stores straight to global memory, compress.cuh:112-154).
 */
#include "compress_kernel.cuh"
#include "runtime.h"

#include <atomic>
#include <cstdlib>

namespace djb {

constexpr uint32_t kCompRowsBig = 8, kCompRowsSmall = 2;

template <uint32_t ROWS, uint32_t STAGES, uint32_t MIN_CTAS, uint32_t COPY = kCopyLsu, bool BASE512 = true, bool PEER = false>
static void launch_compress(cudaStream_t stream, CompressParams &p, Scratch &scratch) {
    constexpr uint32_t TILE = kCompThreads * ROWS * kCompUnit;
    const DeviceProps &dev = device_props();
    auto kernel = compress_kernel<ROWS, STAGES, MIN_CTAS, COPY, BASE512, PEER>;
    constexpr uint32_t smem = compress_smem_bytes<ROWS, STAGES, COPY>();
    // (function attributes and occupancy are per device: one slot per device and instantiation)
    static std::atomic<int> occupancy_of[kMaxDevices] = {};
    int occupancy = occupancy_of[dev.device % kMaxDevices].load(std::memory_order_acquire);
    if (occupancy == 0) {                                   // (idempotent: a race only repeats the queries)
        if (smem > 48 * 1024)
            DJB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        DJB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, kernel, kCompThreads, smem));
        if (occupancy < 1) occupancy = 1;
        occupancy_of[dev.device % kMaxDevices].store(occupancy, std::memory_order_release);
    }
    p.tiles = ceil_div(p.size, TILE);
    const size_t state_bytes = (size_t) p.tiles * 8;
    p.state = (uint64_t *) scratch.device(state_bytes);
    if (p.tiles > 1)        // (a single tile never reads a descriptor)
        DJB_CUDA_CHECK(cudaMemsetAsync(p.state, 0, state_bytes, stream));

    const uint32_t grid = std::min(std::min(p.tiles, dev.sm_count * (uint32_t) occupancy), kCompWindowLoads * kCompThreads);
    if (grid == p.tiles) {
        // one tile per CTA: ordinary launch (see prefix_reduce.cu)
        kernel<<<grid, kCompThreads, smem, stream>>>(p);
    } else {
        // Cooperative launch: all CTAs co-resident (static tile schedule, see scan_kernel.cuh)
        void *args[] = { (void *) &p };
        DJB_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *) kernel, dim3(grid), dim3(kCompThreads), args, smem, stream));
    }
    DJB_POST_LAUNCH();
}

/// Geometry selection + launch. `p` carries the operands (and, for PEER, the communicator view)
template <bool PEER>
static void compress_dispatch(cudaStream_t stream, CompressParams &p, Scratch &scratch) {
    const bool aligned = ((uintptr_t) p.in % 16) == 0;   // TMA source alignment
    // small tiles keep all SMs busy on small masks (cuda_ts.cpp:693 draws its line at 4096)
    const DeviceProps &dev = device_props();
    const bool big = (uint64_t) p.size >= (uint64_t) kCompThreads * kCompRowsBig * kCompUnit * dev.sm_count * 4;
    // copy-out: staging entries written two at a time (kCopyLsuPairs: 0.604 instead of 0.633 ms at 2^30 / 50 %,
    // 0.937 instead of 1.031 ms at 99 %; the bulk shared->global and LDS.128/STG.128 variants lost: profiles/r4c_*, r4e_*)
    if (big && aligned)  launch_compress<kCompRowsBig, 1, 3, kCopyLsuPairs, true, PEER>(stream, p, scratch);
    else if (big)        launch_compress<kCompRowsBig, 0, 3, kCopyLsuPairs, true, PEER>(stream, p, scratch);
    else if (aligned)    launch_compress<kCompRowsSmall, 2, 4, kCopyLsuPairs, true, PEER>(stream, p, scratch);
    else                 launch_compress<kCompRowsSmall, 0, 4, kCopyLsuPairs, true, PEER>(stream, p, scratch);
}

uint32_t compress(cudaStream_t stream, const uint8_t *in, uint32_t size, uint32_t index_base,
                  uint32_t *out, uint32_t *count_dev, bool sync) {
    if (size == 0) { // cuda_ts.cpp:685-686
        if (count_dev)
            DJB_CUDA_CHECK(cudaMemsetAsync(count_dev, 0, sizeof(uint32_t), stream));
        return 0;
    }

    Scratch scratch(stream);
    uint32_t *pinned = scratch.pinned_words();

    CompressParams p{};
    p.in = in; p.out = out; p.size = size; p.index_base = index_base;
    if (count_dev) {
        p.count_out = count_dev;
    } else {
        DJB_CUDA_CHECK(cudaHostGetDevicePointer((void **) &p.count_out, pinned, 0));
    }
    compress_dispatch<false>(stream, p, scratch);

    if (!sync)
        return 0;
    // The reference synchronises here as well (jitc_sync_thread, cuda_ts.cpp:759)
    if (count_dev) {
        DJB_CUDA_CHECK(cudaMemcpyAsync(pinned, count_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    }
    scratch.unlock();                                      // (never block on the GPU with the stream's lock held)
    DJB_CUDA_CHECK(cudaStreamSynchronize(stream));
    return pinned[0];
}

/// An empty shard still takes part in the exchange of the counts
__global__ void peer_counts_kernel(const PeerCtx c, uint32_t mine, uint32_t *host_counts, uint32_t seq) {
    if (threadIdx.x == 0)
        peer_publish_counts(c, mine, host_counts, seq);
}

/// Compaction of this rank's shard of a global mask (indices are global: index_base + local) fused
/// with the exchange of the per-rank counts: counts_host[0..world) (host memory) receives every
/// rank's count, so rank r's list starts at sum(counts_host[0..r)) of the rank-major global list.
/// ONE launch: the thread of the compaction kernel that learns the shard's count exchanges it through
/// the communicator's scalar cells and writes all W counts + a sequence word to pinned host memory;
/// the host spins on that word. Like jit_block_mkperm (cuda_ts.cpp:953-967: table valid, `perm` still
/// being written) the call returns once the counts are known; `out` is complete in stream order.
void comm_compress(cudaStream_t stream, const Comm *comm, const uint8_t *in, uint32_t size, uint32_t index_base,
                   uint32_t *out, uint32_t *counts_host) {
    const uint32_t world = comm_world(comm);
    static std::atomic<uint32_t> next_seq{1};
    uint32_t seq = next_seq.fetch_add(1, std::memory_order_relaxed);
    if (seq == 0) seq = next_seq.fetch_add(1, std::memory_order_relaxed);   // (0 = "nothing published yet")

    Scratch scratch(stream);
    static_assert(Scratch::kPinnedSlotWords >= kMaxPeers + 1, "one pinned word per rank + the sequence word");
    volatile uint32_t *pinned = scratch.pinned_words();
    pinned[kMaxPeers] = 0;

    CompressParams p{};
    p.in = in; p.out = out; p.size = size; p.index_base = index_base;
    p.peer = comm_ctx(comm);
    p.seq = seq;
    DJB_CUDA_CHECK(cudaHostGetDevicePointer((void **) &p.host_counts, (void *) pinned, 0));
    if (size == 0) {
        peer_counts_kernel<<<1, 32, 0, stream>>>(p.peer, 0u, p.host_counts, seq);
        DJB_POST_LAUNCH();
    } else {
        p.count_out = (uint32_t *) scratch.device(256);
        compress_dispatch<true>(stream, p, scratch);
    }
    scratch.unlock();                                      // (never block on the GPU with the stream's lock held)

    // Spin on the sequence word; a stream that has finished (or failed) without publishing it is an error
    for (uint32_t spins = 0; pinned[kMaxPeers] != seq; ++spins) {
        if ((spins & 0xfffu) == 0xfffu) {
            const cudaError_t rv = cudaStreamQuery(stream);
            if (rv != cudaErrorNotReady) {
                DJB_CUDA_CHECK(rv);
                if (pinned[kMaxPeers] != seq)
                    raise(DRJIT_B200_EFATAL, "drjit_b200_comm_compress(): internal error (stream idle, counts missing)");
            }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    for (uint32_t r = 0; r < world; ++r)
        counts_host[r] = pinned[r];
}

} // namespace djb
