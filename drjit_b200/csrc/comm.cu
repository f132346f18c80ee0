/*
 * comm.cu -- peer-memory communicator of one NVSwitch box + the exchange kernels that are not
 * fused into a primitive's own kernel (scalar fold / all-gather of small payloads, all-reduce of
 * a bin array). Protocol and window layout: comm.cuh.
 *
 * The reference has no multi-GPU path. This file never calls NCCL: the
 * windows are mapped once (CUDA IPC between processes, peer access inside one process) and every
 * exchange is a handful of stores and flag spins over NVLink inside a kernel.
 */
#include "comm.cuh"
#include "runtime.h"

#include <cstring>
#include <mutex>

namespace djb {

struct Comm {
    int device = -1;
    uint32_t rank = 0, world = 1;
    size_t bulk_bytes = 0, window_bytes = 0;
    uint8_t *local = nullptr;
    uint8_t *peer[kMaxPeers] = {};
    bool ipc_opened[kMaxPeers] = {};
    bool connected = false;
};

static size_t window_size(size_t bulk_bytes) { return (size_t) kWinBulkOffset + 2 * bulk_bytes; }

PeerCtx comm_ctx(const Comm *c) {
    if (!c)
        raise(DRJIT_B200_EINVAL, "drjit_b200_comm: communicator is NULL!");
    if (!c->connected)
        raise(DRJIT_B200_EINVAL, "drjit_b200_comm: communicator is not connected (call drjit_b200_comm_connect first)!");
    int device = -1;
    DJB_CUDA_CHECK(cudaGetDevice(&device));
    if (device != c->device)
        raise(DRJIT_B200_EINVAL, "drjit_b200_comm: communicator belongs to device %i, but device %i is current!",
              c->device, device);
    PeerCtx ctx{};
    for (uint32_t r = 0; r < kMaxPeers; ++r)
        ctx.win[r] = r < c->world ? c->peer[r] : nullptr;
    ctx.bulk_bytes = c->bulk_bytes;
    ctx.rank = c->rank;
    ctx.world = c->world;
    return ctx;
}

uint32_t comm_rank(const Comm *c) { return c->rank; }
uint32_t comm_world(const Comm *c) { return c->world; }

// ---------------------------------------------------------------------------
//  Scalar fold / small all-gather (one CTA)
// ---------------------------------------------------------------------------
/// dst[0] = fold over the ranks selected by `fold` of every rank's src[0] (type A, op Op)
template <typename Op, typename A>
__global__ void peer_scalar_fold_kernel(const PeerCtx c, const A *src, A *dst, uint32_t fold) {
    if (threadIdx.x == 0)
        *dst = peer_fold_scalar<Op, A>(c, *src, fold);
}

/// dst[r * bytes .. (r + 1) * bytes) = rank r's src[0 .. bytes), bytes % 4 == 0, bytes <= kSlotBytes
__global__ void __launch_bounds__(256) peer_allgather_kernel(const PeerCtx c, const void *src, void *dst, uint32_t bytes) {
    __shared__ uint32_t epoch_smem;
    if (threadIdx.x == 0)
        epoch_smem = peer_begin(c);
    __syncthreads();
    const uint32_t epoch = epoch_smem;
    peer_put_cta(c, epoch, src, bytes);
    peer_wait_cta(c, epoch);
    const uint32_t words = bytes / 4;
    uint32_t *d = reinterpret_cast<uint32_t *>(dst);
    for (uint32_t r = 0; r < c.world; ++r) {
        const uint32_t *s = reinterpret_cast<const uint32_t *>(win_slot(c, c.rank, epoch, r));
        for (uint32_t w = threadIdx.x; w < words; w += blockDim.x)
            d[(size_t) r * words + w] = __ldcg(s + w);
    }
    __syncthreads();
    if (threadIdx.x == 0)
        peer_end(c, epoch);
}

template <typename A> static void scalar_fold_ops(cudaStream_t s, const PeerCtx &c, int op, const void *src, void *dst,
                                                   uint32_t fold, bool is_float) {
    const A *a = (const A *) src; A *d = (A *) dst;
    switch (op) {
        case DRJIT_B200_OP_ADD: peer_scalar_fold_kernel<OpAdd, A><<<1, 32, 0, s>>>(c, a, d, fold); break;
        case DRJIT_B200_OP_MUL: peer_scalar_fold_kernel<OpMul, A><<<1, 32, 0, s>>>(c, a, d, fold); break;
        case DRJIT_B200_OP_MIN: peer_scalar_fold_kernel<OpMin, A><<<1, 32, 0, s>>>(c, a, d, fold); break;
        case DRJIT_B200_OP_MAX: peer_scalar_fold_kernel<OpMax, A><<<1, 32, 0, s>>>(c, a, d, fold); break;
        default:
            if constexpr (std::is_integral<A>::value) {
                if (op == DRJIT_B200_OP_AND) { peer_scalar_fold_kernel<OpAnd, A><<<1, 32, 0, s>>>(c, a, d, fold); break; }
                if (op == DRJIT_B200_OP_OR) { peer_scalar_fold_kernel<OpOr, A><<<1, 32, 0, s>>>(c, a, d, fold); break; }
            }
            (void) is_float;
            raise(DRJIT_B200_EUNSUPPORTED, "drjit_b200_comm_fold(): unsupported reduction type!");
    }
    DJB_POST_LAUNCH();
}

void comm_fold_scalar(cudaStream_t stream, const Comm *comm, int vt, int op, uint32_t fold, const void *src, void *dst) {
    const PeerCtx c = comm_ctx(comm);
    if (fold > kFoldHigher)
        raise(DRJIT_B200_EINVAL, "drjit_b200_comm_fold(): invalid fold mode!");
    switch (vt) {
        case DRJIT_B200_VT_UINT32: scalar_fold_ops<uint32_t>(stream, c, op, src, dst, fold, false); break;
        case DRJIT_B200_VT_INT32:  scalar_fold_ops<int32_t>(stream, c, op, src, dst, fold, false); break;
        case DRJIT_B200_VT_UINT64: scalar_fold_ops<uint64_t>(stream, c, op, src, dst, fold, false); break;
        case DRJIT_B200_VT_INT64:  scalar_fold_ops<int64_t>(stream, c, op, src, dst, fold, false); break;
        case DRJIT_B200_VT_FLOAT32: scalar_fold_ops<float>(stream, c, op, src, dst, fold, true); break;
        case DRJIT_B200_VT_FLOAT64: scalar_fold_ops<double>(stream, c, op, src, dst, fold, true); break;
        default:
            // (u8 / f16 scalars: widen on the caller's side; the fused reductions exchange accumulators)
            raise(DRJIT_B200_EUNSUPPORTED, "drjit_b200_comm_fold(): no kernel for type=%s!", type_name(vt));
    }
}

void comm_allgather(cudaStream_t stream, const Comm *comm, const void *src, uint32_t bytes, void *dst) {
    const PeerCtx c = comm_ctx(comm);
    if (bytes == 0 || bytes % 4 != 0 || bytes > kSlotBytes || ((uintptr_t) src % 4) || ((uintptr_t) dst % 4))
        raise(DRJIT_B200_EINVAL, "drjit_b200_comm_allgather(): payload must be 4-byte aligned, a multiple of 4 "
                                 "and at most %u bytes per rank (got %u)!", kSlotBytes, bytes);
    peer_allgather_kernel<<<1, 256, 0, stream>>>(c, src, dst, bytes);
    DJB_POST_LAUNCH();
}

// ---------------------------------------------------------------------------
//  All-reduce (Add) of an array through the bulk regions: reduce-scatter + all-gather in ONE kernel
// ---------------------------------------------------------------------------
//  n elements, slice r = [r * per, (r + 1) * per) with per a multiple of one 16-byte vector.
//   phase 1: every rank pushes slice p of its array into rank p's staging area [my rank]
//   phase 2: rank r folds the W copies of slice r in rank order (its own copy straight from its
//            array), stores the result into its array and pushes it into every peer's result region
//   phase 3: every rank copies the W - 1 foreign result slices from its result region into its array
//  Grid-wide arrival counters + one flag per (phase, source rank) order the phases across GPUs.
//  The fold order is fixed, and each slice is folded exactly once (by its owner), so all ranks end
//  up with bit-identical arrays. Launched cooperatively: every CTA spins on flags.
constexpr uint32_t kBulkThreads = 512;

template <typename T>
__global__ void __launch_bounds__(kBulkThreads) peer_allreduce_kernel(const PeerCtx c, T *data, uint32_t n, uint32_t per) {
    constexpr uint32_t V = 16 / sizeof(T);
    __shared__ uint32_t last_smem;
    const uint32_t tid = threadIdx.x, W = c.world, me = c.rank;
    const uint32_t epoch = win_header(c, me)->bulk_epoch + 1u;
    uint32_t *counters = win_counters(c);
    const uint64_t gtid = (uint64_t) blockIdx.x * kBulkThreads + tid, gstride = (uint64_t) gridDim.x * kBulkThreads;
    const uint64_t staging_stride = (uint64_t) per * sizeof(T);          // bytes per source rank
    auto slice_len = [&](uint32_t r) -> uint32_t {
        const uint64_t lo = (uint64_t) r * per;
        return lo >= n ? 0u : (uint32_t) min((uint64_t) per, (uint64_t) n - lo);
    };
    // The W - 1 foreign slices move as ONE flat loop over (peer, vector) pairs: every thread of the grid
    // has stores to some peer in flight at the same time (a loop per peer left more than half of the
    // grid idle on every 512 KiB slice and paid one remote latency per peer: 54 us for 4 MB at 8
    // ranks, profiles/r5_n8_time_sharded_n8.txt). A slice's ragged last vector goes element-wise
    // (only the last non-empty slice has one; `data` and all window regions are 16-byte aligned).
    //   dst_of(p), src_of(p): base pointers of slice p's destination / source
    auto copy_foreign = [&](auto dst_of, auto src_of) {
        const uint32_t vps = per / V;                                   // vectors per slice
        const uint64_t total = (uint64_t) (W - 1) * vps;
        for (uint64_t j = gtid; j < total; j += gstride) {
            const uint32_t i = (uint32_t) (j / vps), v = (uint32_t) (j - (uint64_t) i * vps);
            const uint32_t q = (me + 1 + i) % W, len = slice_len(q);
            if (v * V >= len)
                continue;
            T *dst = dst_of(q); const T *src = src_of(q);
            if ((v + 1) * V <= len) {
                reinterpret_cast<uint4 *>(dst)[v] = __ldcg(reinterpret_cast<const uint4 *>(src) + v);
            } else {
                for (uint32_t e = v * V; e < len; ++e)
                    dst[e] = __ldcg(src + e);
            }
        }
    };
    // all CTAs have arrived at counter `which`; the last one raises flag `which + 1` at every peer
    auto grid_arrive_and_flag = [&](uint32_t which) {
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            const uint32_t prev = atomicAdd(counters + which, 1u);
            last_smem = prev == gridDim.x - 1;
        }
        __syncthreads();
        if (last_smem) {
            __threadfence_system();
            if (tid < W && tid != me)
                st_release_sys_u32(win_flags(c, tid, which + 1) + me, epoch);
            if (tid == 0)
                counters[which] = 0;
        }
    };
    auto wait_flags = [&](uint32_t which) {
        if (tid < W && tid != me)
            peer_wait_flag(c, which + 1, tid, epoch);
        __syncthreads();
    };

    // ---- phase 1: push my copy of every foreign slice ------------------------------------------
    copy_foreign([&](uint32_t q) { return reinterpret_cast<T *>(c.win[q] + kWinBulkOffset + me * staging_stride); },
                 [&](uint32_t q) { return (const T *) (data + (uint64_t) q * per); });
    grid_arrive_and_flag(0);
    wait_flags(0);

    // ---- phase 2: fold my slice in rank order, publish the result --------------------------------
    {
        const uint32_t len = slice_len(me);
        T *mine = data + (uint64_t) me * per;
        const uint8_t *staging = c.win[me] + kWinBulkOffset;
        const uint64_t result_off = kWinBulkOffset + c.bulk_bytes + (uint64_t) me * staging_stride;
        const uint32_t nvec = len / V;
        for (uint64_t i = gtid; i < nvec; i += gstride) {
            Vec16<T> acc;
            #pragma unroll
            for (uint32_t e = 0; e < V; ++e) acc.v[e] = T(0);
            for (uint32_t r = 0; r < W; ++r) {
                Vec16<T> x;
                if (r == me) *reinterpret_cast<uint4 *>(&x) = reinterpret_cast<const uint4 *>(mine)[i];
                else *reinterpret_cast<uint4 *>(&x) = __ldcg(reinterpret_cast<const uint4 *>(staging + r * staging_stride) + i);
                #pragma unroll
                for (uint32_t e = 0; e < V; ++e) acc.v[e] = r == 0 ? x.v[e] : acc.v[e] + x.v[e];
            }
            const uint4 raw = *reinterpret_cast<const uint4 *>(&acc);
            reinterpret_cast<uint4 *>(mine)[i] = raw;
            for (uint32_t j = 1; j < W; ++j) {
                const uint32_t p = (me + j) % W;
                reinterpret_cast<uint4 *>(c.win[p] + result_off)[i] = raw;
            }
        }
        for (uint64_t i = (uint64_t) nvec * V + gtid; i < len; i += gstride) {
            T acc = T(0);
            for (uint32_t r = 0; r < W; ++r) {
                const T x = r == me ? mine[i] : __ldcg(reinterpret_cast<const T *>(staging + r * staging_stride) + i);
                acc = r == 0 ? x : acc + x;
            }
            mine[i] = acc;
            for (uint32_t j = 1; j < W; ++j) {
                const uint32_t p = (me + j) % W;
                reinterpret_cast<T *>(c.win[p] + result_off)[i] = acc;
            }
        }
    }
    grid_arrive_and_flag(1);
    wait_flags(1);

    // ---- phase 3: collect the foreign result slices ----------------------------------------------
    copy_foreign([&](uint32_t q) { return data + (uint64_t) q * per; },
                 [&](uint32_t q) { return reinterpret_cast<const T *>(c.win[me] + kWinBulkOffset + c.bulk_bytes + q * staging_stride); });
    __syncthreads();
    if (tid == 0) {
        const uint32_t prev = atomicAdd(counters + 2, 1u);
        if (prev == gridDim.x - 1) {        // the whole grid is done: close the epoch
            counters[2] = 0;
            win_header(c, me)->bulk_epoch = epoch;
        }
    }
}

template <typename T> static void launch_allreduce(cudaStream_t stream, const PeerCtx &c, void *data, uint32_t n) {
    constexpr uint32_t V = 16 / sizeof(T);
    uint32_t per = ceil_div(ceil_div(n, c.world), V) * V;
    if ((uint64_t) per * sizeof(T) * c.world > c.bulk_bytes)
        raise(DRJIT_B200_EINVAL, "drjit_b200_comm_allreduce(): %u elements need %llu bytes of window, the communicator "
                                 "was created with bulk_bytes = %llu!", n,
              (unsigned long long) ((uint64_t) per * sizeof(T) * c.world), (unsigned long long) c.bulk_bytes);
    const DeviceProps &dev = device_props();
    // enough CTAs to drive NVLink (each moves 16 B per thread and round), never more than one per SM
    // (the push / collect phases move W - 1 slices per rank)
    uint32_t grid = std::max(1u, std::min(dev.sm_count, ceil_div((c.world - 1) * (per / V), kBulkThreads)));
    T *d = (T *) data;
    void *args[] = { (void *) &c, (void *) &d, (void *) &n, (void *) &per };
    DJB_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *) peer_allreduce_kernel<T>, dim3(grid), dim3(kBulkThreads),
                                               args, 0, stream));
    DJB_POST_LAUNCH();
}

void comm_allreduce(cudaStream_t stream, const Comm *comm, int vt, int op, void *data, uint32_t n) {
    const PeerCtx c = comm_ctx(comm);
    if (op != DRJIT_B200_OP_ADD)
        raise(DRJIT_B200_EUNSUPPORTED, "drjit_b200_comm_allreduce(): only Add is implemented (scatter-add bins)!");
    if ((uintptr_t) data % 16)
        raise(DRJIT_B200_EINVAL, "drjit_b200_comm_allreduce(): the array must be 16-byte aligned!");
    if (n == 0 || c.world == 1)
        return;
    switch (vt) {
        case DRJIT_B200_VT_FLOAT32: launch_allreduce<float>(stream, c, data, n); break;
        case DRJIT_B200_VT_INT32:
        case DRJIT_B200_VT_UINT32: launch_allreduce<uint32_t>(stream, c, data, n); break;
        case DRJIT_B200_VT_FLOAT64: launch_allreduce<double>(stream, c, data, n); break;
        case DRJIT_B200_VT_INT64:
        case DRJIT_B200_VT_UINT64: launch_allreduce<uint64_t>(stream, c, data, n); break;
        default:
            raise(DRJIT_B200_EUNSUPPORTED, "drjit_b200_comm_allreduce(): no kernel for type=%s!", type_name(vt));
    }
}

// ---------------------------------------------------------------------------
//  Creation / connection
// ---------------------------------------------------------------------------
Comm *comm_create(uint32_t rank, uint32_t world, size_t bulk_bytes) {
    if (world == 0 || world > kMaxPeers || rank >= world)
        raise(DRJIT_B200_EINVAL, "drjit_b200_comm_create(): invalid rank/world (%u/%u; at most %u ranks, one "
                                 "NVSwitch box)!", rank, world, kMaxPeers);
    const DeviceProps &dev = device_props();
    Comm *c = new Comm();
    c->device = dev.device; c->rank = rank; c->world = world;
    c->bulk_bytes = (bulk_bytes + 4095) & ~(size_t) 4095;
    c->window_bytes = window_size(c->bulk_bytes);
    cudaError_t rv = cudaMalloc((void **) &c->local, c->window_bytes);
    if (rv == cudaSuccess)
        rv = cudaMemset(c->local, 0, c->window_bytes);
    if (rv == cudaSuccess)
        rv = cudaDeviceSynchronize();
    if (rv != cudaSuccess) {
        if (c->local) cudaFree(c->local);
        delete c;
        DJB_CUDA_CHECK(rv);
    }
    c->peer[rank] = c->local;
    if (world == 1)
        c->connected = true;
    return c;
}

void comm_handle(const Comm *c, void *handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == DRJIT_B200_COMM_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    DJB_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->local));
    memcpy(handle_out, &h, sizeof(h));
}

void comm_connect_ipc(Comm *c, const void *handles) {
    if (c->connected)
        return;
    const uint8_t *hs = (const uint8_t *) handles;
    for (uint32_t r = 0; r < c->world; ++r) {
        if (r == c->rank)
            continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + (size_t) r * sizeof(h), sizeof(h));
        void *p = nullptr;
        DJB_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[r] = (uint8_t *) p;
        c->ipc_opened[r] = true;
    }
    c->connected = true;
}

void comm_connect_local(Comm **comms, uint32_t world) {
    // every communicator lives in this process: map the windows through plain peer access
    for (uint32_t a = 0; a < world; ++a) {
        Comm *ca = comms[a];
        if (!ca || ca->world != world || ca->rank != a)
            raise(DRJIT_B200_EINVAL, "drjit_b200_comm_connect_local(): communicator %u does not have rank %u of %u!", a, a, world);
    }
    int prev = -1;
    DJB_CUDA_CHECK(cudaGetDevice(&prev));
    for (uint32_t a = 0; a < world; ++a) {
        Comm *ca = comms[a];
        DJB_CUDA_CHECK(cudaSetDevice(ca->device));
        for (uint32_t b = 0; b < world; ++b) {
            Comm *cb = comms[b];
            if (a != b && cb->device != ca->device) {
                int can = 0;
                DJB_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, ca->device, cb->device));
                if (!can) {
                    cudaSetDevice(prev);
                    raise(DRJIT_B200_ECUDA, "drjit_b200_comm_connect_local(): device %i cannot access device %i!",
                          ca->device, cb->device);
                }
                cudaError_t rv = cudaDeviceEnablePeerAccess(cb->device, 0);
                if (rv != cudaSuccess && rv != cudaErrorPeerAccessAlreadyEnabled) {
                    cudaSetDevice(prev);
                    DJB_CUDA_CHECK(rv);
                }
                (void) cudaGetLastError();
            }
            ca->peer[b] = cb->local;
        }
        ca->connected = true;
    }
    DJB_CUDA_CHECK(cudaSetDevice(prev));
}

void comm_destroy(Comm *c) {
    if (!c)
        return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (uint32_t r = 0; r < c->world; ++r)
        if (c->ipc_opened[r])
            cudaIpcCloseMemHandle(c->peer[r]);
    if (c->local)
        cudaFree(c->local);
    if (prev >= 0)
        cudaSetDevice(prev);
    (void) cudaGetLastError();
    delete c;
}

} // namespace djb
