/*
 * comm.cuh -- device side of the multi-GPU combine step.
 *
 * The reference is single-device; sharding the primitives over the GPUs of one NVSwitch box adds
 * exactly one small exchange per primitive (a scalar, W counts, a bucket histogram, a 4 MB bin
 * array). Instead of a library collective + a second kernel, the exchange happens INSIDE the
 * kernel that produces the partial: its last CTA stores the partial into every peer's window over
 * NVLink, raises a flag with release semantics at system scope, waits for the W flags in its own
 * window and folds the W partials in rank order (so every rank computes bit-identical results).
 *
 * Window (one per rank, same layout everywhere, peer-mapped through CUDA IPC or, inside one
 * process, plain peer access):
 *
 *   [Header 256 B][flags 3 x kMaxPeers u32, padded to 256 B][grid counters 256 B]
 *   [scalar cells at 1024: 2 parities x kMaxPeers x 16 B]  <- one-scalar exchanges (reductions, scan totals)
 *   [slots: 2 parities x kMaxPeers x kSlotBytes]          <- small exchanges (<= 64 KiB per rank)
 *   [bulk staging: bulk_bytes][bulk result: bulk_bytes]   <- all-reduce of bin arrays
 *
 * Protocol of a small exchange, epoch e = header.epoch + 1 (device-side counter, so the same
 * kernel arguments work when a CUDA graph replays the call):
 *   put : payload -> win[p].slot[e & 1][rank] for every p, then flag: win[p].flags[rank] = e
 *         (st.release.sys by the thread that wrote the payload, or after a CTA barrier + fence)
 *   wait: spin on own flags[src] >= e for every src (ld.acquire.sys), payloads are then visible
 *   end : header.epoch = e
 * Scalar exchanges (one value of up to 8 bytes per rank) do not use slots and flags: the payload
 * travels inside the flag. A cell is two 8-byte words {epoch << 32 | low half} and {epoch << 32 | high
 * half}; the sender issues the two relaxed 8-byte stores to every peer back to back (8-byte accesses are
 * single-copy atomic, so no fence and no ordering between them is needed) and the receiver polls both
 * words until both carry the epoch. One NVLink latency per exchange instead of one release round trip
 * per peer (8 ranks: 21.7 -> ~6 us per exchange, profiles/r4_n8_time_sharded.txt).
 * Two slot parities suffice: a rank can only begin epoch e + 1 after every peer has flagged e,
 * i.e. after every peer has finished reading the slots of epoch e - 1.
 * All exchanges of one communicator must be enqueued in the same order on every rank and on one
 * stream per rank. A peer that never arrives makes the spin time out (~20 s) and trap: a loud
 * failure instead of a hung GPU.
 */
#pragma once

#include "common.cuh"

namespace djb {

constexpr uint32_t kMaxPeers = 8;                 // one NVSwitch box
constexpr uint32_t kSlotBytes = 64 * 1024;        // payload of one small exchange per rank
constexpr uint32_t kWinHeaderBytes = 256;
constexpr uint32_t kWinFlagsOffset = 256;         // flags[kMaxPeers], bulk_flags1[kMaxPeers], bulk_flags2[kMaxPeers]
constexpr uint32_t kWinCountersOffset = 512;      // grid-arrival counters of the bulk kernel (local use)
constexpr uint32_t kWinCellsOffset = 1024;        // scalar exchanges: [2 parities][kMaxPeers] cells of two u64 words
constexpr uint32_t kWinSlotsOffset = 4096;
constexpr uint64_t kWinBulkOffset = kWinSlotsOffset + 2ull * kMaxPeers * kSlotBytes;

struct WindowHeader {
    uint32_t epoch;          // last completed small exchange
    uint32_t bulk_epoch;     // last completed bulk all-reduce
    uint32_t error;          // set before a time-out trap (diagnostics)
};

/// Kernel-parameter view of a communicator
struct PeerCtx {
    uint8_t *win[kMaxPeers];
    uint64_t bulk_bytes;     // capacity of the staging and of the result region
    uint32_t rank, world;
};

struct Comm;
/// Kernel-parameter view of a connected communicator whose device is current (throws otherwise)
PeerCtx comm_ctx(const Comm *c);

/// How the last CTA folds the gathered partials
enum PeerFold : uint32_t {
    kFoldAll = 0,            // all ranks (all-reduce): every rank obtains the same value
    kFoldLower = 1,          // ranks below mine (exclusive prefix over ranks: scan carries)
    kFoldHigher = 2          // ranks above mine (reverse scans)
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
    uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

__device__ __forceinline__ WindowHeader *win_header(const PeerCtx &c, uint32_t r) {
    return reinterpret_cast<WindowHeader *>(c.win[r]);
}
__device__ __forceinline__ uint32_t *win_flags(const PeerCtx &c, uint32_t r, uint32_t which = 0) {
    return reinterpret_cast<uint32_t *>(c.win[r] + kWinFlagsOffset) + which * kMaxPeers;
}
__device__ __forceinline__ uint32_t *win_counters(const PeerCtx &c) {
    return reinterpret_cast<uint32_t *>(c.win[c.rank] + kWinCountersOffset);
}
__device__ __forceinline__ uint8_t *win_slot(const PeerCtx &c, uint32_t r, uint32_t epoch, uint32_t src) {
    return c.win[r] + kWinSlotsOffset + ((size_t) (epoch & 1u) * kMaxPeers + src) * kSlotBytes;
}

/// Spin until flags[src] of the own window has reached `epoch` (wrap-around safe)
__device__ __forceinline__ void peer_wait_flag(const PeerCtx &c, uint32_t which, uint32_t src, uint32_t epoch) {
    const uint32_t *flag = win_flags(c, c.rank, which) + src;
    if ((int32_t) (ld_acquire_sys_u32(flag) - epoch) >= 0)
        return;
    const uint64_t t0 = global_timer_ns();
    uint32_t spins = 0;
    while ((int32_t) (ld_acquire_sys_u32(flag) - epoch) < 0) {
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 20000000000ull) {
            win_header(c, c.rank)->error = 1000u + src;     // a peer never arrived
            __threadfence_system();
            __trap();
        }
    }
}

// ---- exchanges driven by ONE thread (scalars) ------------------------------------------------
/// Begin the next small exchange: its epoch
__device__ __forceinline__ uint32_t peer_begin(const PeerCtx &c) { return win_header(c, c.rank)->epoch + 1u; }
__device__ __forceinline__ void peer_end(const PeerCtx &c, uint32_t epoch) { win_header(c, c.rank)->epoch = epoch; }

__device__ __forceinline__ void st_relaxed_sys_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_sys_u64(const uint64_t *p) {
    uint64_t v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint64_t *win_cell(const PeerCtx &c, uint32_t r, uint32_t epoch, uint32_t src) {
    return reinterpret_cast<uint64_t *>(c.win[r] + kWinCellsOffset) + ((size_t) (epoch & 1u) * kMaxPeers + src) * 2;
}

/// One thread: publish an 8-byte payload to every rank (including itself). The epoch travels inside
/// both words of the cell: no flag, no fence, all stores in flight together.
__device__ __forceinline__ void peer_put_u64(const PeerCtx &c, uint32_t epoch, uint64_t payload) {
    const uint64_t w0 = ((uint64_t) epoch << 32) | (uint32_t) payload,
                   w1 = ((uint64_t) epoch << 32) | (uint32_t) (payload >> 32);
    for (uint32_t i = 0; i < c.world; ++i) {
        const uint32_t p = (c.rank + i) % c.world;          // start with myself, then round-robin
        uint64_t *cell = win_cell(c, p, epoch, c.rank);
        st_relaxed_sys_u64(cell, w0);
        st_relaxed_sys_u64(cell + 1, w1);
    }
}
/// One thread: the payload rank `src` published in this epoch (spins until both words have arrived)
__device__ __forceinline__ uint64_t peer_get_u64(const PeerCtx &c, uint32_t epoch, uint32_t src) {
    const uint64_t *cell = win_cell(c, c.rank, epoch, src);
    uint64_t w0 = ld_relaxed_sys_u64(cell), w1 = ld_relaxed_sys_u64(cell + 1);
    if ((uint32_t) (w0 >> 32) != epoch || (uint32_t) (w1 >> 32) != epoch) {
        const uint64_t t0 = global_timer_ns();
        uint32_t spins = 0;
        do {
            if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 20000000000ull) {
                win_header(c, c.rank)->error = 1000u + src;     // a peer never arrived
                __threadfence_system();
                __trap();
            }
            w0 = ld_relaxed_sys_u64(cell); w1 = ld_relaxed_sys_u64(cell + 1);
        } while ((uint32_t) (w0 >> 32) != epoch || (uint32_t) (w1 >> 32) != epoch);
    }
    return (w1 << 32) | (uint32_t) w0;
}
/// (the cells need no separate wait: peer_get_u64 spins)
__device__ __forceinline__ void peer_wait_all(const PeerCtx &, uint32_t) { }

/// One thread, whole scalar exchange: returns the fold of the ranks selected by `fold`
template <typename Op, typename A>
__device__ __forceinline__ A peer_fold_scalar(const PeerCtx &c, A mine, uint32_t fold) {
    static_assert(sizeof(A) <= 8, "scalar payloads are at most 8 bytes");
    union { uint64_t u; A a; } pay;
    pay.u = 0; pay.a = mine;
    const uint32_t epoch = peer_begin(c);
    peer_put_u64(c, epoch, pay.u);
    peer_wait_all(c, epoch);
    A acc = Op::template identity<A>();
    const uint32_t lo = fold == kFoldHigher ? c.rank + 1 : 0u,
                   hi = fold == kFoldLower ? c.rank : c.world;
    // Every rank observes EVERY peer's cell of this epoch, also those its fold does not use: the two
    // cell parities are only safe while no rank runs more than one epoch ahead of any other. A rank
    // that skipped the wait (rank 0 of a lower-ranks fold needs no value at all) could publish epoch
    // e + 2 over a cell of epoch e that a slower peer has not read yet -- that peer would then spin
    // for an epoch it can never see (found at world 2 with three scans in a row, profiles/r5_n8*).
    for (uint32_t src = 0; src < c.world; ++src) {
        pay.u = peer_get_u64(c, epoch, src);
        if (src >= lo && src < hi)
            acc = Op::template apply<A>(acc, pay.a);
    }
    peer_end(c, epoch);
    return acc;
}

// ---- exchanges driven by a whole CTA (payloads up to kSlotBytes) -------------------------------
/// All threads of the CTA: copy `bytes` (multiple of 4, source 4-byte aligned) from `src` into
/// every rank's slot for this rank, then raise the flags. Must be followed by peer_wait_cta().
__device__ __forceinline__ void peer_put_cta(const PeerCtx &c, uint32_t epoch, const void *src, uint32_t bytes) {
    const uint32_t words = bytes / 4;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
    for (uint32_t i = 0; i < c.world; ++i) {
        const uint32_t p = (c.rank + i) % c.world;
        uint32_t *d = reinterpret_cast<uint32_t *>(win_slot(c, p, epoch, c.rank));
        for (uint32_t w = threadIdx.x; w < words; w += blockDim.x)
            d[w] = s[w];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < c.world)
        st_release_sys_u32(win_flags(c, (c.rank + threadIdx.x) % c.world) + c.rank, epoch);
}
/// All threads of the CTA: on return every rank's payload of this epoch is readable in the own window
__device__ __forceinline__ void peer_wait_cta(const PeerCtx &c, uint32_t epoch) {
    if (threadIdx.x < c.world)
        peer_wait_flag(c, 0, threadIdx.x, epoch);
    __syncthreads();
}

} // namespace djb
