/*
 * block_reduce.cu -- block reductions, full reductions and dot products.
 *
 * Replaces CUDAThreadState::block_reduce (ext/drjit-core/src/cuda_ts.cpp:195-352) with its
 * 310 precompiled `block_reduce_*` kernels (resources/block_reduce.cuh:83-273), and
 * CUDAThreadState::reduce_dot (cuda_ts.cpp:354-398, resources/reduce_2.cuh:12-76).
 *
 * Design (HBM-bound, read-once):
 *  - "group" kernel for blocks up to 4 KiB: G = 1..32 lanes cooperate on one block, each lane
 *    keeps four 128-bit loads in flight, the result is combined with a G-wide butterfly
 *    (redux.sync for 32-bit integers) -- every block size, vectorised whenever rows are
 *    16-byte aligned (the reference vectorises only block_size >= 1024, cuda_ts.cpp:264-269).
 *  - "chunk" kernel for larger blocks and full reductions: each CTA streams one contiguous
 *    chunk with 4 x 128-bit loads per thread in flight (head/tail peeled, so any element-
 *    aligned pointer is vectorised), writes one partial, and the last CTA to arrive for a
 *    block combines the partials in a fixed order. One launch, no temporary allocation,
 *    deterministic; the reference recurses with 2-3 launches and temp buffers (:347-351).
 */
#include "common.cuh"
#include "comm.cuh"
#include "runtime.h"

namespace djb {

constexpr uint32_t kThreads = 256;

// ---------------------------------------------------------------------------
//  Group kernel
// ---------------------------------------------------------------------------
/// ACROSS: a block needs at most one load per lane of its group (block_size <= G * V); four
/// *blocks* per round then keep four loads in flight per lane (otherwise the four-way unrolling
/// runs along the block and such short blocks would leave a single load in flight).
template <typename T, typename Op, bool VEC, bool ACROSS>
__global__ void __launch_bounds__(kThreads)
block_reduce_group_kernel(const T *__restrict__ in, T *__restrict__ out, uint32_t size,
                          uint32_t block_size, uint32_t block_count, uint32_t log2_g) {
    using A = acc_t<T>;
    constexpr uint32_t V = VEC ? 16 / sizeof(T) : 1;
    const A ident = Op::template identity<A>();

    const uint32_t G = 1u << log2_g,
                   groups_per_cta = kThreads >> log2_g,
                   gl = threadIdx.x & (G - 1),
                   total_groups = gridDim.x * groups_per_cta;

    // Uniform trip count per warp: shuffles below need all 32 lanes
    const uint32_t first_group_of_warp = blockIdx.x * groups_per_cta + ((threadIdx.x & ~31u) >> log2_g);
    const uint32_t my_group_offset = ((threadIdx.x & 31u) >> log2_g);

    if constexpr (ACROSS) {
        for (uint64_t wb = first_group_of_warp; wb < block_count; wb += 4ull * total_groups) {
            A acc[4];
            bool valid[4];
            uint64_t blk[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                blk[u] = wb + (uint64_t) u * total_groups + my_group_offset;
                valid[u] = blk[u] < block_count;
                acc[u] = ident;
                const uint64_t start = blk[u] * block_size, idx = start + (uint64_t) gl * V;
                uint64_t end = start + block_size;
                if (end > size) end = size;
                if (valid[u] && idx < end) {
                    if constexpr (VEC) {
                        const Vec16<T> t = ld_stream<T>(in + idx);
                        acc[u] = to_acc<A>(t.v[0]);
                        #pragma unroll
                        for (uint32_t e = 1; e < V; ++e)
                            acc[u] = Op::template apply<A>(acc[u], to_acc<A>(t.v[e]));
                    } else {
                        acc[u] = to_acc<A>(__ldg(in + idx));
                    }
                }
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (wb + (uint64_t) u * total_groups >= block_count)
                    break;                                  // (uniform per warp)
                for (uint32_t m = G >> 1; m > 0; m >>= 1)
                    acc[u] = Op::template apply<A>(acc[u], shfl_xor(acc[u], m));
                if (gl == 0 && valid[u])
                    out[blk[u]] = from_acc<T>(acc[u]);
            }
        }
    } else {
    for (uint64_t wb = first_group_of_warp; wb < block_count; wb += total_groups) {
        const uint64_t b = wb + my_group_offset;
        const bool valid = b < block_count;
        const uint64_t start = b * block_size;
        uint64_t end = start + block_size;
        if (end > size) end = size;
        if (!valid) end = 0;

        A acc = ident;
        const uint64_t step = (uint64_t) G * V;
        for (uint64_t base = start + (uint64_t) gl * V; base < end; base += 4 * step) {
            if constexpr (VEC) {
                Vec16<T> tmp[4];
                bool ok[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint64_t idx = base + u * step;
                    ok[u] = idx < end;
                    if (ok[u]) tmp[u] = ld_stream<T>(in + idx);
                }
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (ok[u]) {
                        #pragma unroll
                        for (uint32_t e = 0; e < V; ++e)
                            acc = Op::template apply<A>(acc, to_acc<A>(tmp[u].v[e]));
                    }
                }
            } else {
                T tmp[4];
                bool ok[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint64_t idx = base + u * step;
                    ok[u] = idx < end;
                    if (ok[u]) tmp[u] = __ldg(in + idx);
                }
                #pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ok[u]) acc = Op::template apply<A>(acc, to_acc<A>(tmp[u]));
            }
        }

        if (log2_g == 5) {
            acc = WarpReduce<Op, A>::template run<32>(acc);
        } else {
            for (uint32_t m = G >> 1; m > 0; m >>= 1)
                acc = Op::template apply<A>(acc, shfl_xor(acc, m));
        }

        if (gl == 0 && valid)
            out[b] = from_acc<T>(acc);
    }
    }
}

// ---------------------------------------------------------------------------
//  Tiny blocks: block_size elements fill a 16-byte vector, or an integer fraction of one
// ---------------------------------------------------------------------------
//  (dr.block_sum(x, 2 | 4), reductions over a short trailing tensor axis.) The group kernel above
//  gives such a block to one lane with a single load in flight; here every thread streams four
//  128-bit vectors per round (unit stride across lanes), folds the V / BS blocks inside each vector
//  and stores their results as one small vector: block_size 4 on f32 went from 4.2 to ~6 TB/s.
template <typename T, typename Op, uint32_t BS>
__global__ void __launch_bounds__(kThreads)
block_reduce_tiny_kernel(const T *__restrict__ in, T *__restrict__ out, uint64_t n_vec) {
    using A = acc_t<T>;
    constexpr uint32_t V = 16 / sizeof(T), OUTS = V / BS;
    struct alignas(sizeof(T) * OUTS) OutVec { T v[OUTS]; };
    const uint64_t stride = (uint64_t) gridDim.x * kThreads;
    for (uint64_t i = (uint64_t) blockIdx.x * kThreads + threadIdx.x; i < n_vec; i += 4 * stride) {
        Vec16<T> tmp[4];
        bool ok[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            ok[u] = i + u * stride < n_vec;
            if (ok[u]) tmp[u] = ld_stream<T>(in + (i + u * stride) * V);
        }
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            OutVec o;
            #pragma unroll
            for (uint32_t b = 0; b < OUTS; ++b) {
                A acc = to_acc<A>(tmp[u].v[b * BS]);
                #pragma unroll
                for (uint32_t e = 1; e < BS; ++e)
                    acc = Op::template apply<A>(acc, to_acc<A>(tmp[u].v[b * BS + e]));
                o.v[b] = from_acc<T>(acc);
            }
            *reinterpret_cast<OutVec *>(out + (i + u * stride) * OUTS) = o;
        }
    }
}

template <typename T, typename Op, uint32_t BS>
static void launch_tiny(cudaStream_t stream, const T *in, T *out, uint32_t size) {
    constexpr uint32_t V = 16 / sizeof(T);
    const DeviceProps &dev = device_props();
    const uint64_t n_vec = size / V;
    const uint32_t grid = (uint32_t) std::min<uint64_t>((n_vec + 4 * kThreads - 1) / (4 * kThreads), dev.sm_count * 8 * 4);
    block_reduce_tiny_kernel<T, Op, BS><<<std::max(grid, 1u), kThreads, 0, stream>>>(in, out, n_vec);
    DJB_POST_LAUNCH();
}

/// true if the tiny kernel took the call: block_size in {2, 4, 8, 16} dividing the vector width,
/// whole vectors only, aligned input, output aligned to its small vectors
template <typename T, typename Op>
static bool try_tiny(cudaStream_t stream, const T *in, T *out, uint32_t size, uint32_t block_size) {
    constexpr uint32_t V = 16 / sizeof(T);
    if (block_size < 2 || block_size > V || V % block_size || size % V || ((uintptr_t) in % 16) ||
        ((uintptr_t) out % (sizeof(T) * (V / block_size))))
        return false;
    switch (block_size) {
        case 2:  if constexpr (V >= 2)  { launch_tiny<T, Op, 2>(stream, in, out, size);  return true; } break;
        case 4:  if constexpr (V >= 4)  { launch_tiny<T, Op, 4>(stream, in, out, size);  return true; } break;
        case 8:  if constexpr (V >= 8)  { launch_tiny<T, Op, 8>(stream, in, out, size);  return true; } break;
        case 16: if constexpr (V >= 16) { launch_tiny<T, Op, 16>(stream, in, out, size); return true; } break;
    }
    return false;
}

// ---------------------------------------------------------------------------
//  Short blocks that do not divide a vector (block_size 3, 5, 6, 7: sums over a short tensor axis)
// ---------------------------------------------------------------------------
//  A thread owns V consecutive blocks = BS consecutive 128-bit vectors: BS loads in flight, the
//  blocks are folded in registers, the V results leave as one 128-bit store. Neighbouring lanes
//  are BS * 16 bytes apart, so each load instruction touches every sector of the warp's span only
//  partly; the loads are allowed to allocate in L1 so that the other half of a sector is an L1 hit.
template <typename T, typename Op, uint32_t BS>
__global__ void __launch_bounds__(kThreads)
block_reduce_vecblocks_kernel(const T *__restrict__ in, T *__restrict__ out, uint64_t n_groups) {
    using A = acc_t<T>;
    constexpr uint32_t V = 16 / sizeof(T);
    const uint64_t stride = (uint64_t) gridDim.x * kThreads;
    for (uint64_t g = (uint64_t) blockIdx.x * kThreads + threadIdx.x; g < n_groups; g += stride) {
        union { uint4 raw[BS]; T e[BS * V]; } buf;
        const uint4 *src = reinterpret_cast<const uint4 *>(in) + g * BS;
        #pragma unroll
        for (uint32_t k = 0; k < BS; ++k) buf.raw[k] = __ldg(src + k);
        Vec16<T> res;
        #pragma unroll
        for (uint32_t b = 0; b < V; ++b) {
            A acc = to_acc<A>(buf.e[b * BS]);
            #pragma unroll
            for (uint32_t e = 1; e < BS; ++e)
                acc = Op::template apply<A>(acc, to_acc<A>(buf.e[b * BS + e]));
            res.v[b] = from_acc<T>(acc);
        }
        st_stream<T>(out + g * V, res);
    }
}

template <typename T, typename Op, uint32_t BS>
static uint64_t launch_vecblocks(cudaStream_t stream, const T *in, T *out, uint32_t size) {
    constexpr uint32_t V = 16 / sizeof(T);
    const DeviceProps &dev = device_props();
    const uint64_t n_groups = size / (BS * V);
    if (n_groups) {
        const uint32_t grid = (uint32_t) std::min<uint64_t>((n_groups + kThreads - 1) / kThreads, dev.sm_count * 8 * 4);
        block_reduce_vecblocks_kernel<T, Op, BS><<<grid, kThreads, 0, stream>>>(in, out, n_groups);
        DJB_POST_LAUNCH();
    }
    return n_groups * BS * V;       // elements consumed (a whole number of blocks)
}

/// Elements taken by the short-block kernel (0: not applicable); the caller reduces the tail
template <typename T, typename Op>
static uint64_t try_vecblocks(cudaStream_t stream, const T *in, T *out, uint32_t size, uint32_t block_size) {
    if (((uintptr_t) in % 16) || ((uintptr_t) out % 16) || size < (1u << 16))
        return 0;
    switch (block_size) {
        case 3: return launch_vecblocks<T, Op, 3>(stream, in, out, size);
        case 5: return launch_vecblocks<T, Op, 5>(stream, in, out, size);
        case 6: return launch_vecblocks<T, Op, 6>(stream, in, out, size);
        case 7: return launch_vecblocks<T, Op, 7>(stream, in, out, size);
        default: return 0;
    }
}

// ---------------------------------------------------------------------------
//  Chunk kernel (also the dot product when DOT)
// ---------------------------------------------------------------------------
template <typename T, typename Op, bool DOT> struct ChunkAccum {
    using A = acc_t<T>;
    static __device__ __forceinline__ A elem(A acc, T a, T b) {
        if constexpr (DOT) {
            if constexpr (sizeof(A) == 8) return fma((A) a, (A) b, acc);
            else return fmaf(to_acc<A>(a), to_acc<A>(b), acc);
        } else {
            (void) b;
            return Op::template apply<A>(acc, to_acc<A>(a));
        }
    }
};

/// PEER (full reductions of a shard): the CTA that holds the shard's result
/// exchanges it with the other ranks of the communicator through peer memory and folds the W
/// values in rank order before it stores `out` (comm.cuh) -- reduction + combine in one launch.
template <typename T, typename Op, bool DOT, bool VEC, bool PEER = false>
__global__ void __launch_bounds__(kThreads)
block_reduce_chunk_kernel(const T *__restrict__ in, const T *__restrict__ in2, T *__restrict__ out,
                          acc_t<T> *__restrict__ partials, uint32_t *__restrict__ counters,
                          uint32_t size, uint32_t block_size, uint32_t chunk_elems,
                          uint32_t chunks_per_block, const PeerCtx peer = PeerCtx{}, uint32_t fold = 0) {
    using A = acc_t<T>;
    using Acc = ChunkAccum<T, Op, DOT>;
    constexpr uint32_t V = 16 / sizeof(T);
    const A ident = Op::template identity<A>();
    __shared__ A smem[32];
    __shared__ uint32_t is_last_smem;

    const uint32_t block = blockIdx.x / chunks_per_block,
                   chunk = blockIdx.x - block * chunks_per_block;
    const uint64_t block_start = (uint64_t) block * block_size;
    uint64_t start = block_start + (uint64_t) chunk * chunk_elems,
             end = block_start + min((uint64_t) (chunk + 1) * chunk_elems, (uint64_t) block_size);
    if (end > size) end = size;
    if (start > end) start = end;

    A acc = ident;
    const uint32_t tid = threadIdx.x;

    if constexpr (VEC) {
        // Peel scalars up to the first 16-byte boundary, vectorise the body, peel the tail
        const uintptr_t addr = (uintptr_t) (in + start);
        uint64_t head = ((16 - (addr & 15)) & 15) / sizeof(T);
        if (head > end - start) head = end - start;
        const uint64_t body_start = start + head,
                       nvec = (end - body_start) / V,
                       tail_start = body_start + nvec * V;

        if (tid < head)
            acc = Acc::elem(acc, in[start + tid], DOT ? in2[start + tid] : T());
        if (tid < end - tail_start)
            acc = Acc::elem(acc, in[tail_start + tid], DOT ? in2[tail_start + tid] : T());

        const Vec16<T> *vin = reinterpret_cast<const Vec16<T> *>(in + body_start);
        const Vec16<T> *vin2 = reinterpret_cast<const Vec16<T> *>(in2 + body_start);
        for (uint64_t base = tid; base < nvec; base += 4 * kThreads) {
            Vec16<T> ta[4], tb[4];
            bool ok[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t idx = base + u * kThreads;
                ok[u] = idx < nvec;
                if (ok[u]) {
                    ta[u] = ld_stream<T>(vin + idx);
                    if constexpr (DOT) tb[u] = ld_stream<T>(vin2 + idx);
                }
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (ok[u]) {
                    #pragma unroll
                    for (uint32_t e = 0; e < V; ++e)
                        acc = Acc::elem(acc, ta[u].v[e], DOT ? tb[u].v[e] : T());
                }
            }
        }
    } else {
        for (uint64_t base = start + tid; base < end; base += 4 * kThreads) {
            T ta[4], tb[4];
            bool ok[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t idx = base + u * kThreads;
                ok[u] = idx < end;
                if (ok[u]) {
                    ta[u] = __ldg(in + idx);
                    if constexpr (DOT) tb[u] = __ldg(in2 + idx);
                }
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ok[u]) acc = Acc::elem(acc, ta[u], DOT ? tb[u] : T());
        }
    }

    // Partial sums are *added* regardless of Op when computing a dot product
    using Comb = std::conditional_t<DOT, OpAdd, Op>;
    A total = block_reduce<Comb, A, kThreads>(acc, smem, ident);

    if (chunks_per_block == 1) {
        if (tid == 0) {
            if constexpr (PEER) total = peer_fold_scalar<Comb, A>(peer, total, fold);
            out[block] = from_acc<T>(total);
        }
        return;
    }

    // Publish the partial; the last CTA of this block folds all of them in a fixed order
    if (tid == 0) {
        partials[blockIdx.x] = total;
        __threadfence();
        const uint32_t prev = atomicAdd(&counters[block], 1u);
        is_last_smem = prev == chunks_per_block - 1;
    }
    __syncthreads();
    if (!is_last_smem)
        return;
    __threadfence();

    const volatile A *p = partials + (uint64_t) block * chunks_per_block;
    acc = ident;
    for (uint32_t j = tid; j < chunks_per_block; j += kThreads)
        acc = Comb::template apply<A>(acc, p[j]);
    __syncthreads(); // smem reuse
    total = block_reduce<Comb, A, kThreads>(acc, smem, ident);
    if (tid == 0) {
        counters[block] = 0; // leave the control block zeroed for the next call
        if constexpr (PEER) total = peer_fold_scalar<Comb, A>(peer, total, fold);
        out[block] = from_acc<T>(total);
    }
}

// ---------------------------------------------------------------------------
//  Host side
// ---------------------------------------------------------------------------
constexpr uint32_t kGroupMaxBytes = 4096;     // blocks up to this size use the group kernel
constexpr uint32_t kMinChunkBytes = 32768;    // never split a block into chunks smaller than this
constexpr uint32_t kCtasPerSm = 8;

/// Full reduction of a shard + combine over the ranks of a communicator in ONE launch: the chunk
/// kernel over the whole shard (any size, including an empty trailing shard, which contributes the
/// identity), PEER tail.
template <typename T, typename Op>
static void launch_reduce_peer(cudaStream_t stream, const PeerCtx &peer, uint32_t fold, uint32_t size,
                               const T *in, T *out) {
    using A = acc_t<T>;
    constexpr uint32_t V = 16 / sizeof(T);
    const DeviceProps &dev = device_props();
    const uint32_t quantum = kThreads * V * 4;
    const uint64_t bytes = (uint64_t) size * sizeof(T);
    uint32_t cpb = (uint32_t) std::max<uint64_t>(1, std::min<uint64_t>(dev.sm_count * kCtasPerSm, bytes / kMinChunkBytes));
    uint32_t chunk_elems = std::max(quantum, ceil_div(ceil_div(size, cpb), quantum) * quantum);
    cpb = std::max(1u, ceil_div(size, chunk_elems));
    Scratch scratch(stream);
    A *partials = cpb > 1 ? (A *) scratch.device((size_t) cpb * sizeof(A)) : nullptr;
    block_reduce_chunk_kernel<T, Op, false, true, true><<<cpb, kThreads, 0, stream>>>(
        in, nullptr, out, partials, scratch.zeroed_counters(), size, std::max(size, 1u), chunk_elems, cpb, peer, fold);
    DJB_POST_LAUNCH();
}

template <typename T, typename Op>
static void launch_block_reduce(cudaStream_t stream, uint32_t size, uint32_t block_size,
                                const void *in_, void *out_, const PeerCtx *peer = nullptr, uint32_t fold = 0) {
    using A = acc_t<T>;
    const T *in = (const T *) in_;
    T *out = (T *) out_;
    if (peer) {
        launch_reduce_peer<T, Op>(stream, *peer, fold, size, in, out);
        return;
    }
    const DeviceProps &dev = device_props();
    const uint32_t block_count = ceil_div(size, block_size);
    const uint64_t block_bytes = (uint64_t) block_size * sizeof(T);
    constexpr uint32_t V = 16 / sizeof(T);

    if (block_bytes <= 16 && try_tiny<T, Op>(stream, in, out, size, block_size))
        return;
    if (block_size <= 7) {
        const uint64_t done = try_vecblocks<T, Op>(stream, in, out, size, block_size);
        if (done == size)
            return;
        if (done) {     // fewer than V blocks are left: the general path below takes them
            launch_block_reduce<T, Op>(stream, (uint32_t) (size - done), std::min<uint32_t>(block_size, (uint32_t) (size - done)),
                                       in + done, out + done / block_size);
            return;
        }
    }

    if (block_bytes <= kGroupMaxBytes) {
        const bool vec = block_size % V == 0 && size % V == 0 && ((uintptr_t) in % 16) == 0;
        const uint32_t units = vec ? block_size / V : block_size;      // loads per block
        // 2..8 vector loads per block: one load per lane, four blocks per round (ACROSS: block_size
        // 16 on f32 3.9 -> 4.8 TB/s); otherwise up to four loads per lane along the block. (One load
        // per lane was measured slower for element-wise loads and for longer blocks.)
        const bool across = vec && units <= 8;
        uint32_t log2_g = 0;
        while ((1u << log2_g) < 32 && ((across ? 1u : 4u) << log2_g) < units)
            ++log2_g;
        const uint32_t groups_per_cta = kThreads >> log2_g;
        uint32_t grid = ceil_div(block_count, groups_per_cta * (across ? 4u : 1u));
        const uint32_t max_grid = dev.sm_count * kCtasPerSm * 4;
        if (grid > max_grid) grid = max_grid;
        if (grid == 0) grid = 1;
        if (vec && across)
            block_reduce_group_kernel<T, Op, true, true><<<grid, kThreads, 0, stream>>>(
                in, out, size, block_size, block_count, log2_g);
        else if (vec)
            block_reduce_group_kernel<T, Op, true, false><<<grid, kThreads, 0, stream>>>(
                in, out, size, block_size, block_count, log2_g);
        else if (across)
            block_reduce_group_kernel<T, Op, false, true><<<grid, kThreads, 0, stream>>>(
                in, out, size, block_size, block_count, log2_g);
        else
            block_reduce_group_kernel<T, Op, false, false><<<grid, kThreads, 0, stream>>>(
                in, out, size, block_size, block_count, log2_g);
        DJB_POST_LAUNCH();
        return;
    }

    // Chunk kernel: split blocks so that ~kCtasPerSm CTAs per SM are busy
    const uint32_t target_ctas = dev.sm_count * kCtasPerSm;
    uint32_t cpb = 1;
    if (block_count < target_ctas) {
        cpb = ceil_div(target_ctas, block_count);
        const uint32_t max_cpb = (uint32_t) std::max<uint64_t>(1, block_bytes / kMinChunkBytes);
        if (cpb > max_cpb) cpb = max_cpb;
    }
    // chunk size: multiple of a full CTA iteration so that only the last chunk has a ragged end
    const uint32_t quantum = kThreads * V * 4;
    uint32_t chunk_elems = ceil_div(ceil_div(block_size, cpb), quantum) * quantum;
    if (chunk_elems < quantum) chunk_elems = quantum;
    cpb = ceil_div(block_size, chunk_elems);

    Scratch scratch(stream);
    A *partials = nullptr;
    uint32_t *counters = nullptr;
    if (cpb > 1) {
        if (block_count > Scratch::kZeroedCounters)
            raise(DRJIT_B200_EFATAL, "jit_block_reduce(): internal error (counter overflow)");
        partials = (A *) scratch.device((size_t) block_count * cpb * sizeof(A));
        counters = scratch.zeroed_counters();
    }
    const uint64_t grid = (uint64_t) block_count * cpb;
    block_reduce_chunk_kernel<T, Op, false, true><<<(uint32_t) grid, kThreads, 0, stream>>>(
        in, nullptr, out, partials, counters, size, block_size, chunk_elems, cpb);
    DJB_POST_LAUNCH();
}

template <typename T> static void dispatch_op_int(cudaStream_t s, int op, uint32_t size,
                                                  uint32_t bs, const void *in, void *out,
        const PeerCtx *peer = nullptr, uint32_t fold = 0) {
    switch (op) {
        case DRJIT_B200_OP_ADD: launch_block_reduce<T, OpAdd>(s, size, bs, in, out, peer, fold); break;
        case DRJIT_B200_OP_MUL: launch_block_reduce<T, OpMul>(s, size, bs, in, out, peer, fold); break;
        case DRJIT_B200_OP_MIN: launch_block_reduce<T, OpMin>(s, size, bs, in, out, peer, fold); break;
        case DRJIT_B200_OP_MAX: launch_block_reduce<T, OpMax>(s, size, bs, in, out, peer, fold); break;
        case DRJIT_B200_OP_AND: launch_block_reduce<T, OpAnd>(s, size, bs, in, out, peer, fold); break;
        case DRJIT_B200_OP_OR:  launch_block_reduce<T, OpOr>(s, size, bs, in, out, peer, fold); break;
        default: raise(DRJIT_B200_EUNSUPPORTED, "jit_block_reduce(): unsupported reduction type!");
    }
}

template <typename T> static void dispatch_op_minmax(cudaStream_t s, int op, uint32_t size,
                                                     uint32_t bs, const void *in, void *out,
        const PeerCtx *peer = nullptr, uint32_t fold = 0) {
    switch (op) {
        case DRJIT_B200_OP_MIN: launch_block_reduce<T, OpMin>(s, size, bs, in, out, peer, fold); break;
        case DRJIT_B200_OP_MAX: launch_block_reduce<T, OpMax>(s, size, bs, in, out, peer, fold); break;
        default: raise(DRJIT_B200_EFATAL, "jit_block_reduce(): internal dispatch error");
    }
}

template <typename T> static void dispatch_op_float(cudaStream_t s, int vt, int op, uint32_t size,
                                                    uint32_t bs, const void *in, void *out,
        const PeerCtx *peer = nullptr, uint32_t fold = 0) {
    switch (op) {
        case DRJIT_B200_OP_ADD: launch_block_reduce<T, OpAdd>(s, size, bs, in, out, peer, fold); break;
        case DRJIT_B200_OP_MUL: launch_block_reduce<T, OpMul>(s, size, bs, in, out, peer, fold); break;
        case DRJIT_B200_OP_MIN: launch_block_reduce<T, OpMin>(s, size, bs, in, out, peer, fold); break;
        case DRJIT_B200_OP_MAX: launch_block_reduce<T, OpMax>(s, size, bs, in, out, peer, fold); break;
        default:
            // wording of cuda_ts.cpp:313-315
            raise(DRJIT_B200_EUNSUPPORTED,
                  "jit_block_reduce(): no existing kernel for type=%s, op=%s!", type_name(vt),
                  op_name(op));
    }
}

static void dispatch_reduce(cudaStream_t stream, int vt, int op, uint32_t size, uint32_t block_size,
                            const void *in, void *out, const PeerCtx *peer, uint32_t fold) {
    // Signed sum/product/and/or reductions can use the unsigned kernel (cuda_ts.cpp:215-225)
    const bool sign_agnostic = op == DRJIT_B200_OP_ADD || op == DRJIT_B200_OP_MUL ||
                               op == DRJIT_B200_OP_AND || op == DRJIT_B200_OP_OR;
    switch (vt) {
        case DRJIT_B200_VT_BOOL:
        case DRJIT_B200_VT_UINT8:  dispatch_op_int<uint8_t>(stream, op, size, block_size, in, out, peer, fold); break;
        case DRJIT_B200_VT_UINT32: dispatch_op_int<uint32_t>(stream, op, size, block_size, in, out, peer, fold); break;
        case DRJIT_B200_VT_UINT64: dispatch_op_int<uint64_t>(stream, op, size, block_size, in, out, peer, fold); break;
        case DRJIT_B200_VT_INT32:
            if (sign_agnostic) dispatch_op_int<uint32_t>(stream, op, size, block_size, in, out, peer, fold);
            else dispatch_op_minmax<int32_t>(stream, op, size, block_size, in, out, peer, fold);
            break;
        case DRJIT_B200_VT_INT64:
            if (sign_agnostic) dispatch_op_int<uint64_t>(stream, op, size, block_size, in, out, peer, fold);
            else dispatch_op_minmax<int64_t>(stream, op, size, block_size, in, out, peer, fold);
            break;
        case DRJIT_B200_VT_FLOAT16: dispatch_op_float<__half>(stream, vt, op, size, block_size, in, out, peer, fold); break;
        case DRJIT_B200_VT_FLOAT32: dispatch_op_float<float>(stream, vt, op, size, block_size, in, out, peer, fold); break;
        case DRJIT_B200_VT_FLOAT64: dispatch_op_float<double>(stream, vt, op, size, block_size, in, out, peer, fold); break;
        default:
            raise(DRJIT_B200_EUNSUPPORTED, "jit_block_reduce(): no existing kernel for type=%s, op=%s!",
                  type_name(vt), op_name(op));
    }
}


void block_reduce(cudaStream_t stream, int vt, int op, uint32_t size, uint32_t block_size,
                  const void *in, void *out) {
    if (size == 0)
        return;
    if (block_size == 0 || block_size > size) // cuda_ts.cpp:203-207 (reference text says "prefix")
        raise(DRJIT_B200_EINVAL,
              "jit_block_reduce(): invalid block size (size=%u, block_size=%u)!", size, block_size);

    const uint32_t tsize = type_size(vt);
    if (tsize == 0 || op < DRJIT_B200_OP_ADD || op > DRJIT_B200_OP_OR)
        raise(DRJIT_B200_EUNSUPPORTED, "jit_block_reduce(): no existing kernel for type=%s, op=%s!",
              type_name(vt), op_name(op));

    if (block_size == 1) { // cuda_ts.cpp:210-213
        if (in != out)
            DJB_CUDA_CHECK(cudaMemcpyAsync(out, in, (size_t) size * tsize, cudaMemcpyDeviceToDevice, stream));
        return;
    }

    dispatch_reduce(stream, vt, op, size, block_size, in, out, nullptr, 0);
}

/// Full reduction of this rank's shard fused with the combine over all ranks (one launch): every
/// rank's `out[0]` receives the fold selected by `fold` (PeerFold). size == 0 is a valid (empty) shard.
void comm_reduce(cudaStream_t stream, const Comm *comm, int vt, int op, uint32_t fold, uint32_t size,
                 const void *in, void *out) {
    const PeerCtx peer = comm_ctx(comm);
    if (type_size(vt) == 0 || op < DRJIT_B200_OP_ADD || op > DRJIT_B200_OP_OR || fold > kFoldHigher)
        raise(DRJIT_B200_EUNSUPPORTED, "jit_block_reduce(): no existing kernel for type=%s, op=%s!",
              type_name(vt), op_name(op));
    dispatch_reduce(stream, vt, op, size, std::max(size, 1u), in, out, &peer, fold);
}

// ---------------------------------------------------------------------------
//  dr.all / dr.any (src/init.cpp:919-939, src/util.cpp:153-211)
// ---------------------------------------------------------------------------
/// Reduces a bool array with And/Or into four packed bytes. The reference pads the caller's
/// buffer with the identity and reduces it as u32; here the ragged tail is folded in by the
/// kernel itself so nothing is written past values[size).
__global__ void __launch_bounds__(kThreads)
reduce_bool_kernel(const uint8_t *__restrict__ values, uint32_t size, uint32_t *__restrict__ partials,
                   uint32_t *__restrict__ counters, uint8_t *__restrict__ out, int is_and) {
    __shared__ uint32_t smem[32];
    __shared__ uint32_t is_last_smem;
    const uint32_t ident = is_and ? 0xffffffffu : 0u;
    uint32_t acc = ident;

    const uintptr_t addr = (uintptr_t) values;
    uint64_t head = (16 - (addr & 15)) & 15;
    if (head > size) head = size;
    const uint64_t nvec = (size - head) / 16, tail_start = head + nvec * 16;
    const uint64_t gtid = (uint64_t) blockIdx.x * kThreads + threadIdx.x,
                   gstride = (uint64_t) gridDim.x * kThreads;

    auto fold_byte = [&](uint8_t b) {
        // replicate into all four bytes: the position inside the packed word is irrelevant
        const uint32_t w = b * 0x01010101u;
        acc = is_and ? (acc & w) : (acc | w);
    };
    if (gtid < head) fold_byte(values[gtid]);
    if (gtid < size - tail_start) fold_byte(values[tail_start + gtid]);

    const Vec16<uint32_t> *vin = reinterpret_cast<const Vec16<uint32_t> *>(values + head);
    for (uint64_t base = gtid; base < nvec; base += 4 * gstride) {
        Vec16<uint32_t> t[4];
        bool ok[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            ok[u] = base + u * gstride < nvec;
            if (ok[u]) t[u] = ld_stream<uint32_t>(vin + base + u * gstride);
        }
        #pragma unroll
        for (int u = 0; u < 4; ++u)
            if (ok[u]) {
                #pragma unroll
                for (int e = 0; e < 4; ++e)
                    acc = is_and ? (acc & t[u].v[e]) : (acc | t[u].v[e]);
            }
    }

    uint32_t total = is_and ? block_reduce<OpAnd, uint32_t, kThreads>(acc, smem, ident)
                            : block_reduce<OpOr, uint32_t, kThreads>(acc, smem, ident);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = total;
        __threadfence();
        is_last_smem = atomicAdd(counters, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last_smem)
        return;
    __threadfence();
    acc = ident;
    const volatile uint32_t *p = partials;
    for (uint32_t j = threadIdx.x; j < gridDim.x; j += kThreads)
        acc = is_and ? (acc & p[j]) : (acc | p[j]);
    __syncthreads();
    total = is_and ? block_reduce<OpAnd, uint32_t, kThreads>(acc, smem, ident)
                   : block_reduce<OpOr, uint32_t, kThreads>(acc, smem, ident);
    if (threadIdx.x == 0) {
        *reinterpret_cast<uint32_t *>(out) = total; // four packed bools, like the reference's u32 reduce
        *counters = 0;
    }
}

static void launch_reduce_bool(cudaStream_t stream, Scratch &scratch, const uint8_t *values,
                               uint32_t size, uint8_t *out, int op) {
    const DeviceProps &dev = device_props();
    uint32_t grid = ceil_div(size, kThreads * 64);
    grid = std::max(1u, std::min(grid, dev.sm_count * kCtasPerSm));
    uint32_t *partials = (uint32_t *) scratch.device(grid * sizeof(uint32_t));
    reduce_bool_kernel<<<grid, kThreads, 0, stream>>>(values, size, partials, scratch.zeroed_counters(),
                                                      out, op == DRJIT_B200_OP_AND);
    DJB_POST_LAUNCH();
}

void block_reduce_bool(cudaStream_t stream, const uint8_t *values, uint32_t size, uint8_t *out, int op) {
    if (op != DRJIT_B200_OP_AND && op != DRJIT_B200_OP_OR)
        raise(DRJIT_B200_EINVAL, "jit_block_reduce_bool(): op must be And or Or!");
    if (((uintptr_t) out % 4) != 0)
        raise(DRJIT_B200_EINVAL, "jit_block_reduce_bool(): output must be 4-byte aligned!");
    if (size == 0) { // reduction over nothing = identity
        DJB_CUDA_CHECK(cudaMemsetAsync(out, op == DRJIT_B200_OP_AND ? 1 : 0, 4, stream));
        return;
    }
    Scratch scratch(stream);
    launch_reduce_bool(stream, scratch, values, size, out, op);
}

bool all_any(cudaStream_t stream, const uint8_t *values, uint32_t size, int op) {
    if (size == 0)
        return op == DRJIT_B200_OP_AND;
    Scratch scratch(stream);
    uint32_t *pinned = scratch.pinned_words();
    uint8_t *dev_view = nullptr;
    DJB_CUDA_CHECK(cudaHostGetDevicePointer((void **) &dev_view, pinned, 0));
    launch_reduce_bool(stream, scratch, values, size, dev_view, op);
    scratch.unlock();                                      // (never block on the GPU with the stream's lock held)
    DJB_CUDA_CHECK(cudaStreamSynchronize(stream));
    const uint8_t *b = reinterpret_cast<const uint8_t *>(pinned);
    // util.cpp:191,207: combine the four packed partials
    return op == DRJIT_B200_OP_AND ? (b[0] & b[1] & b[2] & b[3]) != 0 : (b[0] | b[1] | b[2] | b[3]) != 0;
}

// ---------------------------------------------------------------------------
//  Dot product
// ---------------------------------------------------------------------------
template <typename T>
static void launch_reduce_dot(cudaStream_t stream, const void *a_, const void *b_, uint32_t size, void *out_,
                              const PeerCtx *peer = nullptr) {
    using A = acc_t<T>;
    constexpr uint32_t V = 16 / sizeof(T);
    const T *a = (const T *) a_, *b = (const T *) b_;
    T *out = (T *) out_;
    const DeviceProps &dev = device_props();

    const uint64_t bytes = (uint64_t) size * sizeof(T);
    uint32_t cpb = dev.sm_count * kCtasPerSm;
    const uint32_t max_cpb = (uint32_t) std::max<uint64_t>(1, bytes / (kMinChunkBytes / 2));
    if (cpb > max_cpb) cpb = max_cpb;
    const uint32_t quantum = kThreads * V * 4;
    uint32_t chunk_elems = std::max(quantum, ceil_div(ceil_div(size, cpb), quantum) * quantum);
    cpb = std::max(1u, ceil_div(size, chunk_elems));

    Scratch scratch(stream);
    A *partials = cpb > 1 ? (A *) scratch.device((size_t) cpb * sizeof(A)) : nullptr;
    uint32_t *counters = scratch.zeroed_counters();
    // both streams must share their misalignment for the 128-bit path
    const bool vec = (((uintptr_t) a ^ (uintptr_t) b) & 15) == 0;
    if (peer && vec)        // shard of a sharded dot product: combine over the ranks in the same launch
        block_reduce_chunk_kernel<T, OpAdd, true, true, true><<<cpb, kThreads, 0, stream>>>(
            a, b, out, partials, counters, size, std::max(size, 1u), chunk_elems, cpb, *peer, kFoldAll);
    else if (peer)
        block_reduce_chunk_kernel<T, OpAdd, true, false, true><<<cpb, kThreads, 0, stream>>>(
            a, b, out, partials, counters, size, std::max(size, 1u), chunk_elems, cpb, *peer, kFoldAll);
    else if (vec)
        block_reduce_chunk_kernel<T, OpAdd, true, true><<<cpb, kThreads, 0, stream>>>(
            a, b, out, partials, counters, size, size, chunk_elems, cpb);
    else
        block_reduce_chunk_kernel<T, OpAdd, true, false><<<cpb, kThreads, 0, stream>>>(
            a, b, out, partials, counters, size, size, chunk_elems, cpb);
    DJB_POST_LAUNCH();
}

void reduce_dot(cudaStream_t stream, int vt, const void *a, const void *b, uint32_t size, void *out) {
    const uint32_t tsize = type_size(vt);
    if (vt != DRJIT_B200_VT_FLOAT16 && vt != DRJIT_B200_VT_FLOAT32 && vt != DRJIT_B200_VT_FLOAT64)
        raise(DRJIT_B200_EUNSUPPORTED, "jit_reduce_dot(): no existing kernel for type=%s!", type_name(vt));
    if (size == 0) { // empty sum
        DJB_CUDA_CHECK(cudaMemsetAsync(out, 0, tsize, stream));
        return;
    }
    switch (vt) {
        case DRJIT_B200_VT_FLOAT16: launch_reduce_dot<__half>(stream, a, b, size, out); break;
        case DRJIT_B200_VT_FLOAT32: launch_reduce_dot<float>(stream, a, b, size, out); break;
        default: launch_reduce_dot<double>(stream, a, b, size, out); break;
    }
}

/// Dot product of this rank's shards fused with the sum over all ranks (one launch)
void comm_reduce_dot(cudaStream_t stream, const Comm *comm, int vt, const void *a, const void *b, uint32_t size,
                     void *out) {
    const PeerCtx peer = comm_ctx(comm);
    switch (vt) {
        case DRJIT_B200_VT_FLOAT16: launch_reduce_dot<__half>(stream, a, b, size, out, &peer); break;
        case DRJIT_B200_VT_FLOAT32: launch_reduce_dot<float>(stream, a, b, size, out, &peer); break;
        case DRJIT_B200_VT_FLOAT64: launch_reduce_dot<double>(stream, a, b, size, out, &peer); break;
        default: raise(DRJIT_B200_EUNSUPPORTED, "jit_reduce_dot(): no existing kernel for type=%s!", type_name(vt));
    }
}

/// dr.all / dr.any over a sharded mask (synchronous like jitc_all/any): local packed flag, scalar
/// And/Or fold over the ranks through peer memory straight into a pinned word, one wait.
bool comm_all_any(cudaStream_t stream, const Comm *comm, const uint8_t *values, uint32_t size, int op) {
    (void) comm_ctx(comm);      // validates
    Scratch scratch(stream);
    uint32_t *pinned = scratch.pinned_words();
    uint32_t *dev_view = nullptr;
    DJB_CUDA_CHECK(cudaHostGetDevicePointer((void **) &dev_view, pinned, 0));
    uint32_t *local = (uint32_t *) scratch.device(256);
    if (size == 0)              // empty trailing shard: identity
        DJB_CUDA_CHECK(cudaMemsetAsync(local, op == DRJIT_B200_OP_AND ? 0xff : 0, 4, stream));
    else
        launch_reduce_bool(stream, scratch, values, size, (uint8_t *) local, op);
    comm_fold_scalar(stream, comm, DRJIT_B200_VT_UINT32, op, kFoldAll, local, dev_view);
    scratch.unlock();
    DJB_CUDA_CHECK(cudaStreamSynchronize(stream));
    const uint8_t *b = reinterpret_cast<const uint8_t *>(pinned);
    return op == DRJIT_B200_OP_AND ? (b[0] & b[1] & b[2] & b[3]) != 0 : (b[0] | b[1] | b[2] | b[3]) != 0;
}

} // namespace djb
