/*
 * prefix_reduce.cu -- (segmented) prefix reductions in a single pass.
 *
 * Replaces CUDAThreadState::block_prefix_reduce (ext/drjit-core/src/cuda_ts.cpp:530-681) and its
 * 280 `block_prefix_reduce_*` kernels (resources/block_prefix_reduce.cuh:33-265).
 *
 * Design: see scan_kernel.cuh (single pass, TMA-staged tiles, early aggregates, decoupled
 * look-back, segmented / reverse / exclusive handled in the same kernel). This file is the host
 * side: argument checks with the reference's wording, type/op dispatch, scratch and launch.
 */
#include "scan_kernel.cuh"
#include "runtime.h"

#include <atomic>
#include <cstdlib>

namespace djb {

// Geometry of the 128-bit path, chosen with scripts/sweep_scan.cu (profiles/r1b_scan_sweep.md):
// 32 KiB tiles, 3 TMA stages, 2 CTAs per SM (192 KiB of tiles in flight per SM), aggregates
// published when a tile lands, windowed carry (scan_kernel.cuh).
constexpr uint32_t kScanRows = 8, kScanStages = 3, kScanCtas = 2;

// ---------------------------------------------------------------------------
//  Host side
// ---------------------------------------------------------------------------
template <typename T, typename Op, bool SEG, bool VEC, uint32_t R, uint32_t STAGES, uint32_t MIN_CTAS>
static void launch_prefix_geom(cudaStream_t stream, PrefixParams &p) {
    using A = acc_t<T>;
    using Geom = ScanGeom<T, VEC, R>;
    const DeviceProps &dev = device_props();
    auto kernel = prefix_reduce_kernel<T, Op, SEG, VEC, R, STAGES, MIN_CTAS>;
    constexpr uint32_t smem = STAGES * Geom::TILE_BYTES;
    constexpr uint32_t threads = ScanRoles<SEG, STAGES>::THREADS;

    // (function attributes and occupancy are per device: one slot per device and instantiation)
    static std::atomic<int> occupancy_of[kMaxDevices] = {};
    int occupancy = occupancy_of[dev.device % kMaxDevices].load(std::memory_order_acquire);
    if (occupancy == 0) {                                   // (idempotent: a race only repeats the queries)
        // (static + dynamic shared memory together may exceed the 48 KiB default)
        if (smem >= 32 * 1024)
            DJB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        DJB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, kernel, threads, smem));
        if (occupancy < 1) occupancy = 1;
        occupancy_of[dev.device % kMaxDevices].store(occupancy, std::memory_order_release);
    }

    // outputs of at least twice the L2 size cannot stay resident anyway (smaller ones are left to the
    // default policy: a consumer kernel finds them in L2)
    // 4-byte types only, where it was measured to help (2^30 u32: 1.363 -> 1.345 ms). The 64-bit scan LOSES a
    // quarter of its speed with these stores (2^28 u64: 0.983 -> 1.263 ms, profiles/r5j_scan_evict_first_ab.txt)
    p.evict_first = sizeof(T) == 4 && (uint64_t) p.size * sizeof(T) >= ((uint64_t) 256 << 20);
#if defined(DJB_AB_NO_EVICT_FIRST)
    p.evict_first = 0;          // (A/B build of the shipped configuration: make AB=... in csrc/Makefile)
#endif
#if defined(DRJIT_B200_EXPERIMENTS)
    if (const char *env = getenv("DRJIT_B200_SCAN_EVICT_FIRST")) p.evict_first = atoi(env) != 0;    // A/B
#endif
    p.tiles = ceil_div(p.size, Geom::TILE);
    Scratch scratch(stream);
    const size_t state_bytes = TileState<A>::bytes(p.tiles);
    p.state = scratch.device(state_bytes);
    // Small arrays (wavefront loops launch these by the thousand): up to four tiles are
    // walked by ONE CTA with the carry in registers -- no descriptor is read, so the memset launch
    // goes away as well (one launch in total; 2^14 u32: 7.4 -> ~6 us, the reference needs 6.2).
    p.single_cta = !SEG && STAGES == 0 && p.tiles > 1 && p.tiles <= 4;
    if (p.tiles > 1 && !p.single_cta)       // (a single tile never reads a descriptor)
        DJB_CUDA_CHECK(cudaMemsetAsync(p.state, 0, state_bytes, stream));

    // (the windowed carry reads one descriptor per CTA of the grid with <= kScanWindowLoads loads per thread)
    const uint32_t grid = p.single_cta ? 1u
        : std::min(std::min(p.tiles, dev.sm_count * (uint32_t) occupancy), kScanWindowLoads * kScanThreads);
    if (grid == p.tiles || p.single_cta) {
        // One tile per CTA: a CTA only ever waits for tiles of lower-numbered CTAs, which were
        // dispatched before it, so an ordinary launch cannot deadlock (and is ~1 us cheaper)
        kernel<<<grid, threads, smem, stream>>>(p);
    } else {
        // Cooperative launch: all CTAs are co-resident, which the static tile schedule relies on
        // (a CTA's window includes tiles that higher-numbered CTAs handled one iteration earlier)
        void *args[] = { (void *) &p };
        DJB_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *) kernel, dim3(grid), dim3(threads), args, smem, stream));
    }
    DJB_POST_LAUNCH();
}

// ---------------------------------------------------------------------------
//  Short blocks (block_size 2..8): every block is scanned inside one thread
// ---------------------------------------------------------------------------
//  dr.block_prefix_sum over a short trailing axis. A thread owns V consecutive blocks = BS
//  consecutive 128-bit vectors (V = elements per vector): BS loads in flight, the scans happen in
//  registers, BS stores. No tiles, no carries, no scratch: the pass runs at copy speed instead of
//  at the 45 % that the per-element head arithmetic of the general segmented kernel allows.
template <typename T, typename Op, uint32_t BS>
__global__ void __launch_bounds__(256)
prefix_short_blocks_kernel(const T *in, T *out, uint64_t n_groups, uint32_t exclusive, uint32_t reverse) {
    using A = acc_t<T>;
    constexpr uint32_t V = 16 / sizeof(T);
    const A ident = Op::template identity<A>();
    const uint64_t stride = (uint64_t) gridDim.x * 256;
    for (uint64_t g = (uint64_t) blockIdx.x * 256 + threadIdx.x; g < n_groups; g += stride) {
        union { uint4 raw[BS]; T e[BS * V]; } buf;
        const uint4 *src = reinterpret_cast<const uint4 *>(in) + g * BS;
        #pragma unroll
        for (uint32_t k = 0; k < BS; ++k) buf.raw[k] = src[k];          // (coherent loads: `out` may alias `in`)
        #pragma unroll
        for (uint32_t b = 0; b < V; ++b) {
            A run = ident;
            #pragma unroll
            for (uint32_t j = 0; j < BS; ++j) {
                const uint32_t i = b * BS + (reverse ? BS - 1 - j : j);
                const A x = to_acc<A>(buf.e[i]);
                const A incl = Op::template apply<A>(run, x);
                buf.e[i] = from_acc<T>(exclusive ? run : incl);
                run = incl;
            }
        }
        uint4 *dst = reinterpret_cast<uint4 *>(out) + g * BS;
        #pragma unroll
        for (uint32_t k = 0; k < BS; ++k) dst[k] = buf.raw[k];
    }
}

/// Elements handled by the short-block kernel (0: not applicable; otherwise a whole number of blocks)
template <typename T, typename Op>
static uint64_t try_prefix_short_blocks(cudaStream_t stream, const PrefixParams &p) {
    constexpr uint32_t V = 16 / sizeof(T);
    const uint32_t bs = p.block_size;
    if (bs < 2 || bs > 8 || p.size < (1u << 16) || p.carry_in || p.total_out ||
        ((uintptr_t) p.in % 16) || ((uintptr_t) p.out % 16))
        return 0;
    const uint64_t n_groups = p.size / (bs * V);
    const uint32_t grid = (uint32_t) std::min<uint64_t>((n_groups + 255) / 256, device_props().sm_count * 32);
    const T *in = (const T *) p.in; T *out = (T *) p.out;
    switch (bs) {
        case 2: prefix_short_blocks_kernel<T, Op, 2><<<grid, 256, 0, stream>>>(in, out, n_groups, p.exclusive, p.reverse); break;
        case 3: prefix_short_blocks_kernel<T, Op, 3><<<grid, 256, 0, stream>>>(in, out, n_groups, p.exclusive, p.reverse); break;
        case 4: prefix_short_blocks_kernel<T, Op, 4><<<grid, 256, 0, stream>>>(in, out, n_groups, p.exclusive, p.reverse); break;
        case 5: prefix_short_blocks_kernel<T, Op, 5><<<grid, 256, 0, stream>>>(in, out, n_groups, p.exclusive, p.reverse); break;
        case 6: prefix_short_blocks_kernel<T, Op, 6><<<grid, 256, 0, stream>>>(in, out, n_groups, p.exclusive, p.reverse); break;
        case 7: prefix_short_blocks_kernel<T, Op, 7><<<grid, 256, 0, stream>>>(in, out, n_groups, p.exclusive, p.reverse); break;
        default: prefix_short_blocks_kernel<T, Op, 8><<<grid, 256, 0, stream>>>(in, out, n_groups, p.exclusive, p.reverse); break;
    }
    DJB_POST_LAUNCH();
    return n_groups * bs * V;
}

template <typename T, typename Op>
static void launch_prefix(cudaStream_t stream, PrefixParams &p) {
    constexpr uint32_t V = 16 / sizeof(T);
    if (const uint64_t done = try_prefix_short_blocks<T, Op>(stream, p)) {
        if (done == p.size)
            return;
        // fewer than V blocks are left (the last one possibly short): the general path takes them.
        // `done` is a multiple of block_size, so the heads of the tail fall where they would in the
        // whole array, in either direction.
        p.in = (const T *) p.in + done;
        p.out = (T *) p.out + done;
        p.size -= (uint32_t) done;
        if (p.block_size > p.size) p.block_size = p.size;
    }
    const bool seg = p.block_size < p.size;
    // 128-bit / TMA path: both pointers 16-byte aligned; mirrored (reverse) vectors additionally
    // need the array end to fall on a vector boundary.
    const bool vec = ((uintptr_t) p.in % 16) == 0 && ((uintptr_t) p.out % 16) == 0 &&
                     (!p.reverse || p.size % V == 0);
#if defined(DRJIT_B200_EXPERIMENTS)
    // A/B of the tile geometry for 8-byte types (scripts/time_prims.py scan64): DRJIT_B200_SCAN64_GEOM=1..3
    if constexpr (sizeof(T) == 8) {
        static const int geom = getenv("DRJIT_B200_SCAN64_GEOM") ? atoi(getenv("DRJIT_B200_SCAN64_GEOM")) : 0;
        if (vec && !seg && geom == 1) return launch_prefix_geom<T, Op, false, true, 4, 3, 4>(stream, p);
        if (vec && !seg && geom == 2) return launch_prefix_geom<T, Op, false, true, 4, 4, 3>(stream, p);
        if (vec && !seg && geom == 3) return launch_prefix_geom<T, Op, false, true, 4, 2, 4>(stream, p);
        if (vec && !seg && geom == 4) return launch_prefix_geom<T, Op, false, true, 4, 3, 3>(stream, p);
    }
#endif
    // Arrays of up to four tiles (32 KiB each): one CTA with direct 128-bit loads -- no TMA ring to
    // fill, no descriptors, no memset: the latency of one launch (wavefront-sized arrays)
    if (vec && !seg && p.size <= 4 * ScanGeom<T, true, kScanRows>::TILE)
        return launch_prefix_geom<T, Op, false, true, kScanRows, 0, kScanCtas>(stream, p);
    if (seg) {
        if (vec) launch_prefix_geom<T, Op, true, true, kScanRows, kScanStages, kScanCtas>(stream, p);
        else     launch_prefix_geom<T, Op, true, false, 4, 0, 1>(stream, p);
    } else {
        if (vec) launch_prefix_geom<T, Op, false, true, kScanRows, kScanStages, kScanCtas>(stream, p);
        else     launch_prefix_geom<T, Op, false, false, 4, 0, 1>(stream, p);
    }
}

template <typename T> static void prefix_ops_int(cudaStream_t s, int op, PrefixParams &p) {
    switch (op) {
        case DRJIT_B200_OP_ADD: launch_prefix<T, OpAdd>(s, p); break;
        case DRJIT_B200_OP_MUL: launch_prefix<T, OpMul>(s, p); break;
        case DRJIT_B200_OP_MIN: launch_prefix<T, OpMin>(s, p); break;
        case DRJIT_B200_OP_MAX: launch_prefix<T, OpMax>(s, p); break;
        case DRJIT_B200_OP_AND: launch_prefix<T, OpAnd>(s, p); break;
        case DRJIT_B200_OP_OR:  launch_prefix<T, OpOr>(s, p); break;
        default: raise(DRJIT_B200_EUNSUPPORTED, "jit_block_prefix_reduce(): unsupported reduction type!");
    }
}
template <typename T> static void prefix_ops_minmax(cudaStream_t s, int op, PrefixParams &p) {
    if (op == DRJIT_B200_OP_MIN) launch_prefix<T, OpMin>(s, p);
    else launch_prefix<T, OpMax>(s, p);
}
template <typename T> static void prefix_ops_float(cudaStream_t s, int vt, int op, PrefixParams &p) {
    switch (op) {
        case DRJIT_B200_OP_ADD: launch_prefix<T, OpAdd>(s, p); break;
        case DRJIT_B200_OP_MUL: launch_prefix<T, OpMul>(s, p); break;
        case DRJIT_B200_OP_MIN: launch_prefix<T, OpMin>(s, p); break;
        case DRJIT_B200_OP_MAX: launch_prefix<T, OpMax>(s, p); break;
        default: // wording of cuda_ts.cpp:638-640
            raise(DRJIT_B200_EUNSUPPORTED,
                  "jit_block_prefix_reduce(): no existing kernel for type=%s, op=%s!",
                  type_name(vt), op_name(op));
    }
}

void block_prefix_reduce(cudaStream_t stream, int vt, int op, uint32_t size, uint32_t block_size,
                         bool exclusive, bool reverse, const void *in, void *out,
                         const void *carry_in, void *total_out) {
    if (size == 0)
        return;
    if (block_size == 0 || block_size > size) // cuda_ts.cpp:540-543
        raise(DRJIT_B200_EINVAL,
              "jit_block_prefix_reduce(): invalid block size (size=%u, block_size=%u)!", size, block_size);

    const uint32_t tsize = type_size(vt);
    if (tsize == 0 || op < DRJIT_B200_OP_ADD || op > DRJIT_B200_OP_OR)
        raise(DRJIT_B200_EUNSUPPORTED, "jit_block_prefix_reduce(): no existing kernel for type=%s, op=%s!",
              type_name(vt), op_name(op));
    const bool flt = vt == DRJIT_B200_VT_FLOAT16 || vt == DRJIT_B200_VT_FLOAT32 || vt == DRJIT_B200_VT_FLOAT64;
    if (flt && (op == DRJIT_B200_OP_AND || op == DRJIT_B200_OP_OR))
        raise(DRJIT_B200_EUNSUPPORTED, "jit_block_prefix_reduce(): no existing kernel for type=%s, op=%s!",
              type_name(vt), op_name(op));

    if (block_size == 1 && !carry_in && !total_out) { // cuda_ts.cpp:544-553
        if (exclusive) {
            const uint64_t ident = reduce_identity(vt, op);
            memset_async(stream, out, size, tsize, &ident);
        } else if (in != out) {
            DJB_CUDA_CHECK(cudaMemcpyAsync(out, in, (size_t) size * tsize, cudaMemcpyDeviceToDevice, stream));
        }
        return;
    }

    // Small arrays: one CTA, every load in flight at once (prefix_small.cu)
    if (block_size == size && prefix_small(stream, vt, op, size, exclusive, reverse, in, out, carry_in, total_out))
        return;
    // Medium blocks: a group of lanes per block (prefix_group.cu). Blocks of 2..8 elements of large
    // arrays stay with the thread-per-block kernel below (copy speed).
    if (block_size < size && !carry_in && !total_out && !(block_size <= 8 && size >= (1u << 16)) &&
        prefix_group_blocks(stream, vt, op, size, block_size, exclusive, reverse, in, out))
        return;

    PrefixParams p{};
    p.in = in; p.out = out; p.size = size; p.block_size = block_size;
    p.exclusive = exclusive; p.reverse = reverse; p.in_place = in == out;
    p.carry_in = carry_in; p.total_out = total_out;
    p.magic = block_size > 1 ? (~0ull / block_size) + 1 : 0; // 2^64 not divisible by bs>1 unless pow2 (then +1 still exact)
    if (block_size > 1 && (block_size & (block_size - 1)) == 0)
        p.magic = (1ull << 63) / (block_size >> 1) + 0; // exact 2^64 / bs; hi-mul gives floor(x/bs) exactly

    const bool sign_agnostic = op == DRJIT_B200_OP_ADD || op == DRJIT_B200_OP_MUL ||
                               op == DRJIT_B200_OP_AND || op == DRJIT_B200_OP_OR;
    switch (vt) {
        case DRJIT_B200_VT_BOOL:
        case DRJIT_B200_VT_UINT8:  prefix_ops_int<uint8_t>(stream, op, p); break;
        case DRJIT_B200_VT_UINT32: prefix_ops_int<uint32_t>(stream, op, p); break;
        case DRJIT_B200_VT_UINT64: prefix_ops_int<uint64_t>(stream, op, p); break;
        case DRJIT_B200_VT_INT32:
            if (sign_agnostic) prefix_ops_int<uint32_t>(stream, op, p);
            else prefix_ops_minmax<int32_t>(stream, op, p);
            break;
        case DRJIT_B200_VT_INT64:
            if (sign_agnostic) prefix_ops_int<uint64_t>(stream, op, p);
            else prefix_ops_minmax<int64_t>(stream, op, p);
            break;
        case DRJIT_B200_VT_FLOAT16: prefix_ops_float<__half>(stream, vt, op, p); break;
        case DRJIT_B200_VT_FLOAT32: prefix_ops_float<float>(stream, vt, op, p); break;
        case DRJIT_B200_VT_FLOAT64: prefix_ops_float<double>(stream, vt, op, p); break;
        default:
            raise(DRJIT_B200_EUNSUPPORTED, "jit_block_prefix_reduce(): no existing kernel for type=%s, op=%s!",
                  type_name(vt), op_name(op));
    }
}

/// Prefix reduction of a global array cut into contiguous per-rank shards.
///  materialise: every element carries the global value. Rank r cannot emit its first output before
///      all lower ranks have read their whole shard, so the minimum for contiguous shards is one
///      extra read pass: launch 1 reduces the shard and exchanges the totals through peer memory
///      inside its last CTA (fold over the lower ranks = this shard's carry), launch 2 is the
///      single-pass scan seeded with that carry. 12 B/element for 4-byte types, no library collective.
///  !materialise ("shard-offset form"): the local scan (8 B/element; its total falls out of the
///      kernel) followed by a one-thread exchange kernel; `out` holds the shard-local prefix and
///      *offset_out the fold over the lower ranks, global[i] = op(offset, local[i]).
/// offset_out (device scalar of type vt, may be NULL when materialising) receives the carry.
void comm_prefix_reduce(cudaStream_t stream, const Comm *comm, int vt, int op, uint32_t size, bool exclusive,
                        bool reverse, const void *in, void *out, void *offset_out, bool materialise) {
    const uint32_t tsize = type_size(vt);
    if (tsize == 0 || op < DRJIT_B200_OP_ADD || op > DRJIT_B200_OP_OR)
        raise(DRJIT_B200_EUNSUPPORTED, "jit_block_prefix_reduce(): no existing kernel for type=%s, op=%s!",
              type_name(vt), op_name(op));
    const uint32_t fold = reverse ? 2u /* ranks above */ : 1u /* ranks below */;
    Scratch scratch(stream);                 // (kept across the inner calls: they nest)
    void *carry = offset_out ? offset_out : scratch.device(256);
    if (materialise) {
        comm_reduce(stream, comm, vt, op, fold, size, in, carry);
        if (size)
            block_prefix_reduce(stream, vt, op, size, size, exclusive, reverse, in, out, carry, nullptr);
        return;
    }
    if (!offset_out)
        raise(DRJIT_B200_EINVAL, "drjit_b200_comm_prefix_reduce(): the shard-offset form needs offset_out!");
    // u8 / f16 totals travel as 4-byte accumulators in the fused reduction; the scalar exchange
    // kernel handles the 4- and 8-byte types the scan totals of this form come in.
    void *total = scratch.device(256);
    if (size) {
        block_prefix_reduce(stream, vt, op, size, size, exclusive, reverse, in, out, nullptr, total);
    } else {
        const uint64_t ident = reduce_identity(vt, op);
        memset_async(stream, total, 1, tsize, &ident);
    }
    comm_fold_scalar(stream, comm, vt, op, fold, total, offset_out);
}

} // namespace djb
