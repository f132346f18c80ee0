/*
 * scan_kernel.cuh -- device side of the single-pass (segmented) prefix reduction.
 * Host-side dispatch: prefix_reduce.cu; geometry sweep: scripts/sweep_scan.cu.
 *
 * One persistent kernel scans the flat array tile by tile. CTA c processes tiles c, c + G,
 * c + 2G, ... (G = grid size); the grid is launched cooperatively, so all CTAs are co-resident.
 * Each thread owns ROWS 128-bit units in a warp-striped arrangement, so global stores are
 * coalesced STG.128 without a shared-memory transpose.
 *
 * STAGES > 0: full tiles are fetched by `cp.async.bulk` (TMA, tma.cuh) into a ring of STAGES
 * shared-memory buffers. A stage is refilled as soon as its tile has been moved to registers,
 * i.e. *before* the tile's scan / store phases, so every CTA keeps STAGES tiles in flight no
 * matter which phase it is in.
 *
 * How a tile learns the reduction of everything before it ("carry"):
 *
 * WINDOW (STAGES >= 2, unsegmented; the 128-bit path of every plain prefix reduction). The eight
 * warps reduce a tile straight from shared memory as soon as its bulk copy has landed -- STAGES-1
 * iterations before they scan it -- and publish the tile's *aggregate*. Aggregates depend on
 * nothing but the input, so there is no chain between CTAs at all: CTA c keeps its own running
 * carry and advances it by the G aggregates of the tiles between its previous tile and the
 * current one,
 *        carry(it) = carry(it-1) (+) agg[tile(it-1)] (+) ... (+) agg[tile(it)-1]
 * (first tile: carry_in (+) agg[0..c)). The G descriptors are loaded by the 256 threads at the top
 * of the iteration (one or two 8-byte loads per thread, L2 hits) and folded into the block-level
 * combine that the scan needs anyway, so their latency hides behind the tile's own load/scan
 * phase and no global round trip is left on the critical path. Every aggregate is read once per
 * CTA: 8 B x G per 32 KiB tile, ~7 % extra L2 (not DRAM) traffic. Measured against the
 * decoupled look-back that this replaces (scripts/sweep_scan.cu, profiles/r1b_scan_sweep.md):
 * the look-back cost 0.85 us per tile that two CTAs per SM could not hide (1.69 ms vs 1.31 ms
 * with the chain disabled for 2^30 u32); the windowed carry removes it. A side effect: the
 * association order of a floating-point prefix sum is fixed by (size, grid), so results are
 * reproducible run to run, which a look-back (whose window depends on timing) cannot offer.
 *
 * LOOK-BACK (segmented scans, the element-wise path for unaligned arrays): decoupled look-back
 * (Merrill & Garland) by warp 0, 320 predecessor descriptors per round with all loads of a round
 * in flight. `block_size` only changes where the running value is reset: the scan is segmented
 * with heads at multiples of `block_size`; a tile that contains a head publishes its post-head
 * aggregate as a complete prefix immediately, so short blocks never form a dependency chain.
 *
 * `reverse` mirrors tile and element order; `exclusive` shifts the result by one element at
 * store time.
 *
 * Reference: resources/block_prefix_reduce.cuh:46-214 (one element per thread, 10-step
 * Hillis-Steele scan with 20 barriers per 1024 elements, every warp spins in the look-back).
 */
#pragma once

#include "common.cuh"
#include "tma.cuh"

namespace djb {

constexpr uint32_t kScanThreads = 256;
constexpr uint32_t kScanWarps = kScanThreads / 32;
constexpr uint32_t kScanFetchTid = kScanThreads - 32;   // lane 0 of the last scan warp issues the TMA copies
constexpr uint32_t kLookbackLoads = 10;                 // descriptors per lane and round (window = 320)

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

/// 8-byte accumulators: {value, status} in one 16-byte word, written and read with a single
/// 128-bit transaction (STG.E.128.STRONG.GPU / LDG.E.128.STRONG.GPU). A naturally aligned
/// 16-byte access is performed as one transaction by the memory system -- the same property
/// CUB's ScanTileState relies on for 8-byte values -- so a reader sees either the old or the
/// new {value, status} pair, never a mix, and no acquire/release pair (and no second, dependent
/// load) is needed.
template <typename A> struct TileState<A, 8> {
    ulonglong2 *words;
    static size_t bytes(uint32_t tiles) { return (size_t) tiles * 16; }
    __host__ __device__ void bind(void *base, uint32_t) { words = (ulonglong2 *) base; }
    __device__ __forceinline__ void publish(uint32_t tile, uint32_t status, A value) {
        uint64_t bits;
        memcpy(&bits, &value, 8);
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};"
                     :: "l"(words + tile), "l"(bits), "l"((uint64_t) status) : "memory");
    }
    __device__ __forceinline__ void load(uint32_t tile, uint32_t &status, A &value) {
        uint64_t bits, st;
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];"
                     : "=l"(bits), "=l"(st) : "l"(words + tile) : "memory");
        status = (uint32_t) st;
        memcpy(&value, &bits, 8);
    }
};

struct PrefixParams {
    const void *in;
    void *out;
    void *state;            // tile descriptors (zero on entry)
    const void *carry_in;   // optional device scalar (sharded scans)
    void *total_out;        // optional device scalar
    uint64_t magic;         // floor(2^64 / block_size) (+ 1 unless block_size is a power of two)
    uint32_t size, block_size, tiles;
    uint8_t exclusive, reverse, in_place;
    uint8_t evict_first;    // output much larger than L2: store with the evict-first policy (st.global.cs), so that
                            // the kernel leaves fewer dirty lines behind for its successor to share bandwidth
                            // with (bench step: scan 1.402 -> 1.390 ms, the compress after it 0.658 -> 0.648 ms)
    uint8_t single_cta;     // grid of one CTA walking all tiles in order: the carry stays in registers, the
                            // tile descriptors are neither read nor expected to be zero (no memset launch)
    uint8_t debug;          // read only in -DDRJIT_B200_EXPERIMENTS builds (scripts/sweep_scan.cu): 1 = skip look-back, 2 = skip stores
};

/// Tile geometry. A "unit" is what one thread moves at once (a 128-bit vector, or one element
/// on the unaligned path); a thread owns ROWS units. R = units per thread for 4-byte types;
/// narrower types use fewer units so that the number of accumulator registers (elements per
/// thread) stays the same.
template <typename T, bool VEC, uint32_t R> struct ScanGeom {
    static constexpr uint32_t V = VEC ? 16 / sizeof(T) : 1;
    static constexpr uint32_t ROWS_ = !VEC ? R : (V >= 16 ? R / 4 : (V == 8 ? R / 2 : R));
    static constexpr uint32_t ROWS = ROWS_ == 0 ? 1 : ROWS_;
    static constexpr uint32_t TILE = kScanThreads * ROWS * V;          // elements
    static constexpr uint32_t TILE_BYTES = TILE * sizeof(T);
};

/// Scans with at least two TMA stages publish aggregates early and use the windowed carry
template <bool SEG, uint32_t STAGES> struct ScanRoles {
    static constexpr bool WINDOW = STAGES >= 2;
    static constexpr uint32_t THREADS = kScanThreads;
};
constexpr uint32_t kScanWindowLoads = 3;    // descriptors per thread: grids of up to 768 CTAs

/// Barrier among the eight scan warps
__device__ __forceinline__ void scan_warps_sync() {
    asm volatile("bar.sync 1, %0;" :: "n"(kScanThreads) : "memory");
}

/// Decoupled look-back executed by one full warp: reduction of all tiles before `tile`
/// (tile >= 1). Each round inspects kLookbackLoads x 32 predecessors with all descriptor loads
/// of the round in flight together; windows of 32 are then folded nearest first, and only a
/// window that still holds an unpublished descriptor is polled again.
template <typename Op, typename A>
__device__ __forceinline__ A scan_lookback(TileState<A> &state, uint32_t tile, uint32_t lane) {
    const A ident = Op::template identity<A>();
    A excl = ident;
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
        uint32_t status[kLookbackLoads];
        A value[kLookbackLoads];
        #pragma unroll
        for (uint32_t j = 0; j < kLookbackLoads; ++j) {
            status[j] = kPrefix; value[j] = ident;
            const int32_t idx = pred - 32 * (int32_t) j;
            if (idx >= 0) state.load((uint32_t) idx, status[j], value[j]);
        }
        bool found = false;
        #pragma unroll
        for (uint32_t j = 0; j < kLookbackLoads; ++j) {
            if (!found && consume(pred - 32 * (int32_t) j, status[j], value[j]))
                found = true;
        }
        if (found) break;
        pred -= 32 * (int32_t) kLookbackLoads;
    }
    return excl;
}

template <typename T, typename Op, bool SEG, bool VEC, uint32_t R, uint32_t STAGES, uint32_t MIN_CTAS>
__global__ void __launch_bounds__(kScanThreads, MIN_CTAS)
prefix_reduce_kernel(const PrefixParams p) {
    using A = acc_t<T>;
    using Geom = ScanGeom<T, VEC, R>;
    constexpr uint32_t V = Geom::V;         // elements per unit
    constexpr uint32_t ROWS = Geom::ROWS;   // units per thread
    constexpr uint32_t TILE = Geom::TILE;
    constexpr bool STAGED = STAGES > 0;
    constexpr bool WINDOW = ScanRoles<SEG, STAGES>::WINDOW;
    constexpr uint32_t NS = STAGED ? STAGES : 1;
    // (the one-CTA walk of small arrays is only ever launched on the unstaged instantiation: keep its
    //  branches out of the TMA kernel)
    constexpr bool SINGLE = !SEG && !STAGED;
    static_assert(!STAGED || VEC, "staged tiles need the 128-bit path");
    const A ident = Op::template identity<A>();

    extern __shared__ __align__(128) uint8_t stage_mem[];   // STAGES x TILE_BYTES
    __shared__ uint64_t full_bar[NS];       // TMA: tile has landed in its stage
    __shared__ A early_val[NS][kScanWarps]; // WINDOW: per-warp partial aggregates of a landed tile
    __shared__ A warp_val[kScanWarps];
    __shared__ A win_val[kScanWarps];       // WINDOW: per-warp partial sums of the carry window
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

    // SEG: largest head position (scan order) <= x, or -1 if there is none. Forward scans have
    // heads at multiples of block_size; mirrored scans where (size - position) is one.
    auto last_head_le = [&](uint64_t x) -> int64_t {
        const uint32_t d = rev ? mod_bs((uint64_t) size - x) : mod_bs(x);
        const uint32_t k = rev ? (d == 0 ? 0u : bs - d) : d;
        return (int64_t) x - (int64_t) k;
    };

    // ---- tile acquisition (static round-robin schedule) ---------------------------------
    uint64_t policy = 0;
    auto tile_of = [&](uint32_t k) -> uint64_t { return (uint64_t) blockIdx.x + (uint64_t) k * gridDim.x; };
    auto tile_is_staged = [&](uint64_t tile) -> bool {
        return STAGED && tile < p.tiles && (tile + 1) * TILE <= size;
    };
    // fetch thread only: start the bulk copy of `tile` into stage s
    auto issue = [&](uint32_t s, uint64_t tile) {
        if (tile_is_staged(tile)) {
            const uint64_t first = rev ? (uint64_t) size - (tile + 1) * TILE : tile * TILE;
            mbar_expect_tx(&full_bar[s], Geom::TILE_BYTES);
            bulk_load(stage_mem + (size_t) s * Geom::TILE_BYTES, in + first, Geom::TILE_BYTES, &full_bar[s], policy);
        }
    };
    if constexpr (STAGED) {
        if (tid == kScanFetchTid) {
            #pragma unroll
            for (uint32_t s = 0; s < STAGES; ++s)
                mbar_init(&full_bar[s], 1);
            fence_proxy_async();
            policy = policy_evict_first();
            #pragma unroll
            for (uint32_t s = 0; s < STAGES; ++s) issue(s, tile_of(s));
        }
        __syncthreads();
    }

    // WINDOW: per-warp partial aggregate of this CTA's k-th tile, read from its stage as soon as
    // it has landed; after the next barrier warp 0 folds the partials and publishes the aggregate
    auto early_reduce = [&](uint32_t k) {
        const uint64_t t64 = tile_of(k);
        if (!tile_is_staged(t64))
            return;
        const uint32_t s = k % NS;
        mbar_wait(&full_bar[s], (k / NS) & 1u);
        const uint8_t *src = stage_mem + (size_t) s * Geom::TILE_BYTES;
        // SEG: the aggregate only covers what follows the tile's last head (tile-local position)
        uint32_t first = 0;
        if constexpr (SEG) {
            const int64_t h = last_head_le((t64 + 1) * TILE - 1);
            if (h > (int64_t) (t64 * TILE)) first = (uint32_t) (h - (int64_t) (t64 * TILE));
        }
        A acc = ident;
        #pragma unroll
        for (uint32_t r = 0; r < ROWS; ++r) {
            const uint32_t u = (warp * ROWS + r) * 32 + lane;       // unit in scan order
            if (SEG && (u + 1) * V <= first)
                continue;                                           // entirely before the tile's last head
            Vec16<T> v;
            *reinterpret_cast<uint4 *>(&v) = lds128(src + (SEG && rev ? Geom::TILE_BYTES - (u + 1) * 16 : u * 16));
            #pragma unroll
            for (uint32_t e = 0; e < V; ++e) {
                const A x = to_acc<A>(SEG && rev ? v.v[V - 1 - e] : v.v[e]);
                if (!SEG || u * V + e >= first)
                    acc = Op::template apply<A>(acc, x);
            }
        }
        acc = WarpReduce<Op, A>::template run<32>(acc);
        if (lane == 0)
            early_val[s][warp] = acc;
    };
    auto early_publish = [&](uint32_t k) {
        const uint64_t t64 = tile_of(k);
        if (!tile_is_staged(t64))
            return;
        A v = lane < kScanWarps ? early_val[k % NS][lane] : ident;
        v = WarpReduce<Op, A>::template run<32>(v);
        if (lane == 0)
            state.publish((uint32_t) t64, kAggregate, v);
    };
    if constexpr (WINDOW) {
        for (uint32_t k = 0; k + 1 < STAGES; ++k) {
            early_reduce(k);
            scan_warps_sync();
            if (warp == 0) early_publish(k);
        }
    }

    A carry = ident;            // WINDOW: reduction of everything before the current tile
    A prev_total = ident;       // single_cta: aggregate of the tile handled in the previous iteration
    if (WINDOW || (SINGLE && p.single_cta)) {
        if (p.carry_in) carry = to_acc<A>(*(const T *) p.carry_in);
    }

    for (uint32_t it = 0;; ++it) {
        const uint64_t tile64 = tile_of(it);
        if (tile64 >= p.tiles)
            break;
        const uint32_t tile = (uint32_t) tile64;
        const uint32_t stage = it % NS;
        const uint64_t tile_base = tile64 * TILE;   // scan-order position
        const bool staged = tile_is_staged(tile64);

        // ---- WINDOW: start loading the aggregates between the previous tile and this one ----
        // (published at least STAGES-1 iterations ago; consumed before the combine barrier)
        uint32_t win_status[kScanWindowLoads];
        A win_value[kScanWindowLoads];
        uint32_t win_lo = it == 0 ? 0u : tile - gridDim.x;
        bool win_reset = false;     // SEG: a head lies inside the window, the running carry restarts there
        if constexpr (SEG && WINDOW) {
            if (tile != 0) {
                const int64_t h = last_head_le(tile_base - 1);
                if (h >= 0 && (uint32_t) ((uint64_t) h / TILE) >= win_lo) {
                    win_lo = (uint32_t) ((uint64_t) h / TILE);
                    win_reset = true;
                }
            }
        }
        const uint32_t win_n = ((DJB_DEBUG(p.debug) & 1) || (SINGLE && p.single_cta)) ? 0u : tile - win_lo;   // (debug: carry chain disabled)
        if constexpr (WINDOW) {
            #pragma unroll
            for (uint32_t j = 0; j < kScanWindowLoads; ++j) {
                win_status[j] = kAggregate; win_value[j] = ident;
                const uint32_t o = j * kScanThreads + tid;
                if (o < win_n) state.load(win_lo + o, win_status[j], win_value[j]);
            }
        }

        // ---- load + thread-local segmented scan ----------------------------------
        A incl[ROWS][V];
        uint32_t head_mask[ROWS];   // bit e: element e of the unit starts a segment
        A unit_val[ROWS];           // aggregate after the last head of the unit

        Vec16<T> raw[ROWS];
        if constexpr (WINDOW)
            early_reduce(it + STAGES - 1);
        if constexpr (STAGED) {
            if (staged) {
                mbar_wait(&full_bar[stage], (it / NS) & 1u);
                const uint8_t *src = stage_mem + (size_t) stage * Geom::TILE_BYTES;
                #pragma unroll
                for (uint32_t k = 0; k < ROWS; ++k) {
                    const uint32_t u = (warp * ROWS + k) * 32 + lane;
                    const uint32_t off = rev ? Geom::TILE_BYTES - (u + 1) * 16 : u * 16;
                    *reinterpret_cast<uint4 *>(&raw[k]) = lds128(src + off);
                }
            }
            scan_warps_sync();                       // stage is free again
            if (tid == kScanFetchTid)                // refill before the phases below
                issue(stage, tile64 + (uint64_t) STAGES * gridDim.x);
            if constexpr (WINDOW) {
                if (warp == 0) early_publish(it + STAGES - 1);
            }
        }

        uint32_t res = 0, row_step = 0;     // SEG: residue bookkeeping of this thread's units
        if constexpr (SEG) {
            const uint64_t first = tile_base + (uint64_t) ((warp * ROWS * 32 + lane) * V);
            if (first < size)
                res = rev ? mod_bs((uint64_t) size - first) : mod_bs(first);
            row_step = mod_bs(32 * V);
        }
        #pragma unroll
        for (uint32_t k = 0; k < ROWS; ++k) {
            const uint64_t s0 = tile_base + (uint64_t) (((warp * ROWS + k) * 32 + lane) * V);
            A x[V];
            if (staged) {
                #pragma unroll
                for (uint32_t e = 0; e < V; ++e)
                    x[e] = to_acc<A>(rev ? raw[k].v[V - 1 - e] : raw[k].v[e]);
            } else if (s0 >= size) {
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
                // forward: head iff i % bs == 0; reverse: head iff (i + 1) % bs == 0, i = size-1-s.
                // `res` is the residue of the unit's first element (stepped from row to row).
                if (s0 < size) {
                    if (bs >= V) {          // at most one head per unit
                        const uint32_t d = rev ? res : (res == 0 ? 0u : bs - res);
                        hm = d < V ? 1u << d : 0u;
                    } else {
                        uint32_t r = res;
                        #pragma unroll
                        for (uint32_t e = 0; e < V; ++e) {
                            hm |= (r == 0 ? 1u : 0u) << e;
                            if (rev) r = r == 0 ? bs - 1 : r - 1;
                            else     r = r + 1 == bs ? 0 : r + 1;
                        }
                    }
                }
                if (rev) res = res >= row_step ? res - row_step : res + bs - row_step;
                else     { res += row_step; if (res >= bs) res -= bs; }
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
        if constexpr (WINDOW) {
            // fold this thread's share of the carry window (an aggregate that is not there yet
            // belongs to a CTA that runs behind: poll)
            A w = ident;
            #pragma unroll
            for (uint32_t j = 0; j < kScanWindowLoads; ++j) {
                const uint32_t o = j * kScanThreads + tid;
                while (win_status[j] == kInvalid) {
                    __nanosleep(20);
                    state.load(win_lo + o, win_status[j], win_value[j]);
                }
                w = Op::template apply<A>(w, win_value[j]);
            }
            w = WarpReduce<Op, A>::template run<32>(w);
            if (lane == 0) win_val[warp] = w;
        }
        scan_warps_sync();

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

        // ---- the tile's carry ----------------------------------------------------------
        A tile_carry;
        if constexpr (WINDOW) {
            if (SEG && win_reset) carry = ident;
            #pragma unroll
            for (uint32_t w = 0; w < kScanWarps; ++w)
                carry = Op::template apply<A>(carry, win_val[w]);
            if (SINGLE && p.single_cta && it > 0)           // one CTA: the previous tile was mine
                carry = Op::template apply<A>(carry, prev_total);
            prev_total = tv;
            tile_carry = carry;
            if (!staged && tid == 0)     // (a ragged last tile was not published early; keeps the
                state.publish(tile, kAggregate, tv);  //  descriptor array fully defined)
        } else if (SINGLE && p.single_cta) {
            // one CTA walks all tiles in order: every thread knows the tile aggregate, the running
            // value stays in a register (no descriptor, no look-back)
            tile_carry = carry;
            carry = Op::template apply<A>(carry, tv);
            // (the next tile rewrites warp_val / warp_flag: every warp must have read them. The unstaged
            //  path has no other barrier between this read and that write -- racecheck found the hazard,
            //  profiles/r5n_racecheck_single_cta.txt)
            scan_warps_sync();
        } else {
            // decoupled look-back by warp 0 while the other warps wait
            if (warp == 0) {
                A excl = ident;
                if (tile == 0) {
                    if (p.carry_in) excl = to_acc<A>(*(const T *) p.carry_in);
                    if (lane == 0)
                        state.publish(0, kPrefix, tf ? tv : Op::template apply<A>(excl, tv));
                } else if (DJB_DEBUG(p.debug) & 1) {
                    excl = ident;
                } else {
                    if (lane == 0 && tf)
                        state.publish(tile, kPrefix, tv);      // complete: the segment starts inside
                    else if (lane == 0)
                        state.publish(tile, kAggregate, tv);
                    excl = scan_lookback<Op, A>(state, tile, lane);
                    if (lane == 0 && !tf)
                        state.publish(tile, kPrefix, Op::template apply<A>(excl, tv));
                }
                if (lane == 0)
                    carry_smem = excl;
            }
            scan_warps_sync();
            tile_carry = carry_smem;
        }
        if (p.total_out && tile == p.tiles - 1 && tid == 0)
            *(T *) p.total_out = from_acc<T>(tf ? tv : Op::template apply<A>(tile_carry, tv));

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
                if ((DJB_DEBUG(p.debug) & 2) && v.v[0] != T(12345))
                    continue;
                if (p.evict_first) {
                    const uint4 r = *reinterpret_cast<const uint4 *>(&v);
                    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
                                 :: "l"(rev ? out + (size - s0 - V) : out + s0), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
                    continue;
                }
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

} // namespace djb
