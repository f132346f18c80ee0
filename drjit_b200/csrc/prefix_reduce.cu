/*
 * prefix_reduce.cu -- (segmented) prefix reductions in a single pass.
 *
 * Replaces CUDAThreadState::block_prefix_reduce (ext/drjit-core/src/cuda_ts.cpp:530-681) and its
 * 280 `block_prefix_reduce_*` kernels (resources/block_prefix_reduce.cuh:33-265).
 *
 * Design. One persistent kernel scans the flat array tile by tile (4096 elements per tile for
 * types up to 4 bytes, 2048 for 8-byte types): each thread owns 16 (8) elements held as
 * 128-bit vectors in a warp-striped arrangement, so loads and stores are fully coalesced
 * LDG.128/STG.128 without a shared-memory transpose. Tiles are handed out through an atomic
 * ticket (forward progress for the look-back does not depend on block scheduling order) and
 * are chained with a decoupled look-back (Merrill & Garland) in which a whole warp inspects
 * 32 predecessor descriptors per step. `block_size` only changes where the running value is
 * reset: the scan is *segmented*, with segment heads at multiples of `block_size`; a tile
 * that contains a head publishes its post-head aggregate as a complete prefix immediately,
 * so short blocks never form a dependency chain. `reverse` mirrors the tile order and the
 * element order; `exclusive` shifts the result by one element at store time.
 *
 * The reference processes one element per thread with a 10-step Hillis-Steele scan in shared
 * memory (20 barriers per 1024 elements) and lets every warp of every CTA spin in the
 * look-back (block_prefix_reduce.cuh:133-202).
 */
#include "common.cuh"
#include "runtime.h"

#include <cstdlib>

namespace djb {

constexpr uint32_t kScanThreads = 256;
constexpr uint32_t kScanWarps = kScanThreads / 32;

enum : uint32_t { kInvalid = 0, kAggregate = 1, kPrefix = 2 };

// ---------------------------------------------------------------------------
//  Tile descriptors
// ---------------------------------------------------------------------------
template <typename A, size_t Size = sizeof(A)> struct TileState;

/// 4-byte accumulators: {value, status} packed into one 64-bit word (single-copy atomic)
template <typename A> struct TileState<A, 4> {
    uint64_t *words;
    static size_t bytes(uint32_t tiles) { return (size_t) tiles * 8; }
    __host__ __device__ void bind(void *base, uint32_t) { words = (uint64_t *) base; }
    __device__ __forceinline__ void publish(uint32_t tile, uint32_t status, A value) {
        uint32_t bits;
        memcpy(&bits, &value, 4);
        st_relaxed_u64(words + tile, ((uint64_t) bits << 32) | status);
    }
    __device__ __forceinline__ void load(uint32_t tile, uint32_t &status, A &value) {
        const uint64_t w = ld_relaxed_u64(words + tile);
        status = (uint32_t) w;
        const uint32_t bits = (uint32_t) (w >> 32);
        memcpy(&value, &bits, 4);
    }
};

/// 8-byte accumulators: separate value arrays guarded by a status word (release/acquire)
template <typename A> struct TileState<A, 8> {
    uint32_t *status_words;
    uint64_t *aggregates, *prefixes;
    static size_t bytes(uint32_t tiles) { return ((size_t) tiles * 4 + 255) / 256 * 256 + (size_t) tiles * 16; }
    __host__ __device__ void bind(void *base, uint32_t tiles) {
        status_words = (uint32_t *) base;
        aggregates = (uint64_t *) ((uint8_t *) base + ((size_t) tiles * 4 + 255) / 256 * 256);
        prefixes = aggregates + tiles;
    }
    __device__ __forceinline__ void publish(uint32_t tile, uint32_t status, A value) {
        uint64_t bits;
        memcpy(&bits, &value, 8);
        st_relaxed_u64((status == kPrefix ? prefixes : aggregates) + tile, bits);
        st_release_u32(status_words + tile, status);
    }
    __device__ __forceinline__ void load(uint32_t tile, uint32_t &status, A &value) {
        status = ld_acquire_u32(status_words + tile);
        uint64_t bits = 0;
        if (status != kInvalid)
            bits = ld_relaxed_u64((status == kPrefix ? prefixes : aggregates) + tile);
        memcpy(&value, &bits, 8);
    }
};

struct PrefixParams {
    const void *in;
    void *out;
    void *state;            // tile descriptors
    uint32_t *ticket;       // zero on entry
    const void *carry_in;   // optional device scalar (sharded scans)
    void *total_out;        // optional device scalar
    uint64_t magic;         // floor(2^64 / block_size) + 1
    uint32_t size, block_size, tiles;
    uint8_t exclusive, reverse, in_place;
};

// ---------------------------------------------------------------------------
//  Kernel
// ---------------------------------------------------------------------------
/// Tile geometry. A "unit" is what one thread loads at once (a 128-bit vector, or one element
/// on the unaligned path); a thread owns ROWS units. Large arrays use 4x larger tiles: at
/// B200 bandwidth a 16 KiB tile would start ~200 tiles/us, more than a 64-descriptor
/// look-back window can follow ("prefix_reduce").
template <typename T, bool VEC, bool BIG> struct ScanGeom {
    static constexpr uint32_t V = VEC ? 16 / sizeof(T) : 1;
    static constexpr uint32_t BASE_ROWS = VEC ? (V >= 16 ? 1 : (V == 8 ? 2 : 4)) : 4;
    static constexpr uint32_t ROWS = BASE_ROWS * (BIG ? 4 : 1);
    static constexpr uint32_t TILE = kScanThreads * ROWS * V;
};

template <typename T, typename Op, bool SEG, bool VEC, bool BIG>
__global__ void __launch_bounds__(kScanThreads)
prefix_reduce_kernel(const PrefixParams p) {
    using A = acc_t<T>;
    using Geom = ScanGeom<T, VEC, BIG>;
    constexpr uint32_t V = Geom::V;         // elements per unit
    constexpr uint32_t ROWS = Geom::ROWS;   // units per thread
    constexpr uint32_t TILE = Geom::TILE;
    const A ident = Op::template identity<A>();

    __shared__ uint32_t tile_smem;
    __shared__ A warp_val[kScanWarps];
    __shared__ uint32_t warp_flag[kScanWarps];
    __shared__ A carry_smem;

    const T *in = (const T *) p.in;
    T *out = (T *) p.out;
    TileState<A> state;
    state.bind(p.state, p.tiles);

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t size = p.size, bs = p.block_size;
    const bool rev = p.reverse;

    // residue helper: x mod block_size for x <= 2^32 (exact)
    auto mod_bs = [&](uint64_t x) -> uint32_t {
        const uint64_t q = __umul64hi(x, p.magic);
        return (uint32_t) (x - q * bs);
    };

    while (true) {
        if (tid == 0)
            tile_smem = atomicAdd(p.ticket, 1u);
        __syncthreads();
        const uint32_t tile = tile_smem;
        if (tile >= p.tiles)
            break;
        const uint64_t tile_base = (uint64_t) tile * TILE;   // scan-order position

        // ---- load + thread-local segmented scan ----------------------------------
        A incl[ROWS][V];
        uint32_t head_mask[ROWS];   // bit e: element e of the unit starts a segment
        A unit_val[ROWS];           // aggregate after the last head of the unit
        #pragma unroll
        for (uint32_t k = 0; k < ROWS; ++k) {
            const uint64_t s0 = tile_base + (uint64_t) (((warp * ROWS + k) * 32 + lane) * V);
            A x[V];
            if (s0 >= size) {
                #pragma unroll
                for (uint32_t e = 0; e < V; ++e) x[e] = ident;
            } else if (VEC && s0 + V <= size) {
                const T *src = rev ? in + (size - s0 - V) : in + s0;
                Vec16<T> v = p.in_place ? ld_vec<T>(src) : ld_stream<T>(src);
                #pragma unroll
                for (uint32_t e = 0; e < V; ++e)
                    x[e] = to_acc<A>(rev ? v.v[V - 1 - e] : v.v[e]);
            } else {
                #pragma unroll
                for (uint32_t e = 0; e < V; ++e) {
                    const uint64_t s = s0 + e;
                    x[e] = s < size ? to_acc<A>(in[rev ? size - 1 - s : s]) : ident;
                }
            }

            uint32_t hm = 0;
            if constexpr (SEG) {
                if (s0 < size) {
                    // forward: head iff i % bs == 0; reverse: head iff (i + 1) % bs == 0, i = size-1-s
                    uint32_t r = rev ? mod_bs((uint64_t) size - s0) : mod_bs(s0);
                    #pragma unroll
                    for (uint32_t e = 0; e < V; ++e) {
                        hm |= (r == 0 ? 1u : 0u) << e;
                        if (rev) r = r == 0 ? bs - 1 : r - 1;
                        else     r = r + 1 == bs ? 0 : r + 1;
                    }
                }
            }
            head_mask[k] = hm;

            A run = ident;
            #pragma unroll
            for (uint32_t e = 0; e < V; ++e) {
                if (SEG && ((hm >> e) & 1u)) run = x[e];
                else run = Op::template apply<A>(run, x[e]);
                incl[k][e] = run;
            }
            unit_val[k] = run;
        }

        // ---- warp-level: scan the units of each row across lanes, chain the rows ---
        A unit_prefix[ROWS];        // value entering the unit, from inside this warp
        uint32_t unit_pflag = 0;    // bit k: a head precedes unit k inside this warp
        A wcarry = ident;
        bool wflag = false;
        #pragma unroll
        for (uint32_t k = 0; k < ROWS; ++k) {
            uint32_t hb = 0, seg = 0;
            if constexpr (SEG) {
                hb = __ballot_sync(kFullMask, head_mask[k] != 0);
                const uint32_t le = hb & lanemask_le();
                seg = le ? 31u - __clz(le) : 0u;
            }
            A v = unit_val[k];
            #pragma unroll
            for (uint32_t d = 1; d < 32; d <<= 1) {
                const A t = shfl_up(v, d);
                if (lane >= d + seg)
                    v = Op::template apply<A>(t, v);
            }
            A ex = shfl_up(v, 1);
            if (lane == 0) ex = ident;
            const bool ef = SEG && (hb & lanemask_lt()) != 0;
            unit_prefix[k] = ef ? ex : Op::template apply<A>(wcarry, ex);
            if (wflag || ef) unit_pflag |= 1u << k;

            const A row_val = shfl_idx(v, 31);
            const bool row_flag = SEG && hb != 0;
            wcarry = row_flag ? row_val : Op::template apply<A>(wcarry, row_val);
            wflag = wflag || row_flag;
        }
        if (lane == 0) {
            warp_val[warp] = wcarry;
            warp_flag[warp] = wflag;
        }
        __syncthreads();

        // ---- CTA-level: prefix over the preceding warps, tile aggregate -----------
        A pv = ident, tv = ident;
        bool pf = false, tf = false;
        #pragma unroll
        for (uint32_t w = 0; w < kScanWarps; ++w) {
            const A wv = warp_val[w];
            const bool wf = SEG && warp_flag[w];
            if (w == warp) { pv = tv; pf = tf; }
            tv = wf ? wv : Op::template apply<A>(tv, wv);
            tf = tf || wf;
        }

        // ---- decoupled look-back (warp 0) -----------------------------------------
        if (warp == 0) {
            A excl = ident;
            if (tile == 0) {
                if (p.carry_in) excl = to_acc<A>(*(const T *) p.carry_in);
                if (lane == 0)
                    state.publish(0, kPrefix, tf ? tv : Op::template apply<A>(excl, tv));
            } else {
                if (lane == 0 && tf)
                    state.publish(tile, kPrefix, tv);      // complete: the segment starts inside
                else if (lane == 0)
                    state.publish(tile, kAggregate, tv);

                // Each round inspects 64 predecessors (two descriptors per lane, both loads in
                // flight together): window 0 = the nearest 32 tiles, window 1 = the 32 before.
                int32_t pred = (int32_t) tile - 1 - (int32_t) lane;
                // folds one window of 32 descriptors into `excl`; true once a complete prefix was found
                auto consume = [&](int32_t first, uint32_t status, A value) -> bool {
                    while (__any_sync(kFullMask, status == kInvalid)) {
                        __nanosleep(20);
                        if (first >= 0)
                            state.load((uint32_t) first, status, value);
                    }
                    const uint32_t done = __ballot_sync(kFullMask, status == kPrefix);
                    // nearest predecessor holding a complete prefix (lowest lane)
                    const uint32_t stop = done ? (uint32_t) __ffs(done) - 1 : 31u;
                    A contrib = lane <= stop ? value : ident;
                    contrib = WarpReduce<Op, A>::template run<32>(contrib);
                    excl = Op::template apply<A>(contrib, excl);
                    return done != 0;
                };
                while (true) {
                    // lanes past the start of the array act like a finished tile holding the identity
                    uint32_t status0 = kPrefix, status1 = kPrefix;
                    A value0 = ident, value1 = ident;
                    if (pred >= 0) state.load((uint32_t) pred, status0, value0);
                    if (pred >= 32) state.load((uint32_t) (pred - 32), status1, value1);
                    if (consume(pred, status0, value0)) break;
                    if (consume(pred - 32, status1, value1)) break;
                    pred -= 64;
                }
                if (lane == 0 && !tf)
                    state.publish(tile, kPrefix, Op::template apply<A>(excl, tv));
            }
            if (lane == 0) {
                carry_smem = excl;
                if (p.total_out && tile == p.tiles - 1)
                    *(T *) p.total_out = from_acc<T>(tf ? tv : Op::template apply<A>(excl, tv));
            }
        }
        __syncthreads();
        const A tile_carry = carry_smem;

        // ---- combine and store -----------------------------------------------------
        const A warp_in = pf ? pv : Op::template apply<A>(tile_carry, pv);
        #pragma unroll
        for (uint32_t k = 0; k < ROWS; ++k) {
            const uint64_t s0 = tile_base + (uint64_t) (((warp * ROWS + k) * 32 + lane) * V);
            if (s0 >= size)
                continue;
            const bool cut = (unit_pflag >> k) & 1u;
            const A enter = cut ? unit_prefix[k] : Op::template apply<A>(warp_in, unit_prefix[k]);
            const uint32_t hm = head_mask[k];

            A res[V];
            bool seen = false;
            A prev = enter;                         // inclusive value of the previous element
            #pragma unroll
            for (uint32_t e = 0; e < V; ++e) {
                const bool head = SEG && ((hm >> e) & 1u);
                seen = seen || head;
                const A inc = seen ? incl[k][e] : Op::template apply<A>(enter, incl[k][e]);
                res[e] = p.exclusive ? (head ? ident : prev) : inc;
                prev = inc;
            }

            if (VEC && s0 + V <= size) {
                Vec16<T> v;
                #pragma unroll
                for (uint32_t e = 0; e < V; ++e)
                    v.v[rev ? V - 1 - e : e] = from_acc<T>(res[e]);
                st_stream<T>(rev ? out + (size - s0 - V) : out + s0, v);
            } else {
                #pragma unroll
                for (uint32_t e = 0; e < V; ++e) {
                    const uint64_t s = s0 + e;
                    if (s < size)
                        out[rev ? size - 1 - s : s] = from_acc<T>(res[e]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
//  Host side
// ---------------------------------------------------------------------------
template <typename T, typename Op, bool SEG, bool VEC, bool BIG>
static void launch_prefix_geom(cudaStream_t stream, PrefixParams &p) {
    using A = acc_t<T>;
    constexpr uint32_t TILE = ScanGeom<T, VEC, BIG>::TILE;
    const DeviceProps &dev = device_props();

    static int occupancy = 0; // per instantiation
    if (occupancy == 0) {
        DJB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occupancy, prefix_reduce_kernel<T, Op, SEG, VEC, BIG>, kScanThreads, 0));
        if (occupancy < 1) occupancy = 1;
    }

    p.tiles = ceil_div(p.size, TILE);
    Scratch scratch(stream);
    const size_t state_bytes = TileState<A>::bytes(p.tiles);
    uint8_t *mem = (uint8_t *) scratch.device(256 + state_bytes);
    p.ticket = (uint32_t *) mem;
    p.state = mem + 256;
    DJB_CUDA_CHECK(cudaMemsetAsync(mem, 0, 256 + state_bytes, stream));

    const uint32_t grid = std::min(p.tiles, dev.sm_count * (uint32_t) occupancy);
    prefix_reduce_kernel<T, Op, SEG, VEC, BIG><<<grid, kScanThreads, 0, stream>>>(p);
    DJB_POST_LAUNCH();
}

/// Small tiles keep every SM busy on small arrays; large tiles bound the tile rate on big ones
static int scan_big_override() {
    static int v = -2;
    if (v == -2) {
        const char *env = getenv("DRJIT_B200_SCAN_BIG"); // developer override: 0 / 1
        v = env ? atoi(env) : -1;
    }
    return v;
}

template <typename T, typename Op, bool SEG, bool VEC>
static void launch_prefix_variant(cudaStream_t stream, PrefixParams &p) {
    const DeviceProps &dev = device_props();
    bool big = (uint64_t) p.size >= (uint64_t) ScanGeom<T, VEC, true>::TILE * dev.sm_count * 8;
    if (scan_big_override() >= 0) big = scan_big_override() != 0;
    if (big) launch_prefix_geom<T, Op, SEG, VEC, true>(stream, p);
    else     launch_prefix_geom<T, Op, SEG, VEC, false>(stream, p);
}

template <typename T, typename Op>
static void launch_prefix(cudaStream_t stream, PrefixParams &p) {
    constexpr uint32_t V = 16 / sizeof(T);
    const bool seg = p.block_size < p.size;
    // 128-bit path: both pointers 16-byte aligned; mirrored (reverse) vectors additionally
    // need the array end to fall on a vector boundary.
    const bool vec = ((uintptr_t) p.in % 16) == 0 && ((uintptr_t) p.out % 16) == 0 &&
                     (!p.reverse || p.size % V == 0);
    if (seg) {
        if (vec) launch_prefix_variant<T, Op, true, true>(stream, p);
        else     launch_prefix_variant<T, Op, true, false>(stream, p);
    } else {
        if (vec) launch_prefix_variant<T, Op, false, true>(stream, p);
        else     launch_prefix_variant<T, Op, false, false>(stream, p);
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

/// Reduction identity as raw bits (jitc_reduce_identity, src/var.cpp:2642-2652)
static uint64_t reduce_identity(int vt, int op) {
    const uint32_t ts = type_size(vt);
    const bool sgn = vt == DRJIT_B200_VT_INT8 || vt == DRJIT_B200_VT_INT16 ||
                     vt == DRJIT_B200_VT_INT32 || vt == DRJIT_B200_VT_INT64;
    const bool flt = vt == DRJIT_B200_VT_FLOAT16 || vt == DRJIT_B200_VT_FLOAT32 || vt == DRJIT_B200_VT_FLOAT64;
    const uint64_t ones = ts == 8 ? ~0ull : ((1ull << (8 * ts)) - 1);
    switch (op) {
        case DRJIT_B200_OP_AND: return ones;
        case DRJIT_B200_OP_MUL:
            if (!flt) return 1;
            return vt == DRJIT_B200_VT_FLOAT16 ? 0x3C00ull : vt == DRJIT_B200_VT_FLOAT32 ? 0x3F800000ull : 0x3FF0000000000000ull;
        case DRJIT_B200_OP_MIN:
            if (flt) return vt == DRJIT_B200_VT_FLOAT16 ? 0x7C00ull : vt == DRJIT_B200_VT_FLOAT32 ? 0x7F800000ull : 0x7FF0000000000000ull;
            return sgn ? ones >> 1 : ones;
        case DRJIT_B200_OP_MAX:
            if (flt) return vt == DRJIT_B200_VT_FLOAT16 ? 0xFC00ull : vt == DRJIT_B200_VT_FLOAT32 ? 0xFF800000ull : 0xFFF0000000000000ull;
            return sgn ? (ones >> 1) + 1 : 0;
        default: return 0; // Add, Or
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

} // namespace djb
