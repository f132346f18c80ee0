/*
 * runtime.h -- host-side plumbing shared by all primitives: error handling, per-device
 * properties, scratch arenas, pinned staging words, launch accounting.
 *
 * Replaces, for this path only, the pieces of the reference that sit between
 * CUDAThreadState and the driver: device table (src/cuda_core.cpp:266-539),
 * submit_gpu() (src/cuda_ts.cpp:12-47) and the scratch use of jitc_malloc()
 * (src/malloc.cpp:102-200).
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdexcept>
#include <string>

#include "../../include/drjit_b200.h"

namespace djb {

/// Error carrying one of the DRJIT_B200_E* codes; caught at the C-ABI boundary (api.cu)
struct Error : std::runtime_error {
    int code;
    Error(int code, const std::string &msg) : std::runtime_error(msg), code(code) {}
};

[[noreturn]] void raise(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

#define DJB_CUDA_CHECK(expr)                                                             \
    do {                                                                                 \
        cudaError_t err__ = (expr);                                                      \
        if (err__ != cudaSuccess) {                                                      \
            (void) cudaGetLastError(); /* do not leak a launch error into the next call */ \
            ::djb::raise(DRJIT_B200_ECUDA, "cuda_check(): API error %04i (%s): \"%s\" in " \
                         "%s:%i.", (int) err__, cudaGetErrorName(err__),                 \
                         cudaGetErrorString(err__), __FILE__, __LINE__);                 \
        }                                                                                \
    } while (0)

constexpr int kMaxDevices = 64;     // per-device caches of function attributes / occupancy

struct DeviceProps {
    int device = -1;
    uint32_t sm_count = 0;
    uint32_t smem_optin = 0;     // max opt-in dynamic shared memory per CTA
    uint32_t cc = 0;             // compute capability * 10
};

/// Properties of the device current on this thread (cached; throws ECUDA if unusable)
const DeviceProps &device_props();

/// Scratch memory for one primitive call. Obtained from the user allocator hooks when
/// installed (jitc_malloc-style, released when the object dies -- the hook's free must be
/// stream-ordered like jitc_free(), see drjit_b200.h), otherwise from a grow-only arena owned
/// by the library and keyed by (device, stream): successive kernels on one stream are ordered,
/// so the arena can be reused without a sync. The arena never shrinks and is never freed while
/// work may still use it: a grown arena's predecessor is retired behind an event. While the
/// stream is being captured into a CUDA graph the arena cannot grow (cudaMalloc is not
/// capturable): reserve it beforehand with drjit_b200_reserve_scratch() or one warm-up call.
///
/// The object holds the stream's mutex from construction until unlock() / destruction, i.e.
/// while kernels are being *enqueued*; it must be unlock()ed before the host blocks on the GPU.
class Scratch {
public:
    Scratch(cudaStream_t stream);
    ~Scratch();
    /// Release the stream's mutex early (before cudaStreamSynchronize / cudaEventSynchronize).
    /// No device() / reserve() calls afterwards; pointers handed out stay valid for the kernels
    /// already enqueued, and pinned_slot() stays private to this call.
    void unlock();
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;

    /// Make sure the arena can serve `total_bytes` (sum of the 256-byte-rounded sizes
    /// of all later device() calls). Must precede the first device() call of a
    /// primitive that needs more than one buffer.
    void reserve(size_t total_bytes);
    /// Device memory, 256-byte aligned, valid until the Scratch object is destroyed
    void *device(size_t bytes);
    /// The first `zeroed_bytes()` of the stream's control block: always zero on entry to a
    /// primitive; kernels that use it must leave it zeroed (self-cleaning counters).
    uint32_t *zeroed_counters();
    static constexpr size_t kZeroedCounters = 4096; // number of u32 counters
    /// kPinnedSlotWords pinned, device-mapped host words (results read by the host after a wait).
    /// Slots rotate per call (kPinnedSlots of them per stream), so a call that has unlock()ed can
    /// still read its own slot after the wait while another thread enqueues on the same stream.
    uint32_t *pinned_words();
    static constexpr size_t kPinnedSlots = 32, kPinnedSlotWords = 16, kPinnedWords = kPinnedSlotWords * kPinnedSlots;

    struct StreamState;

private:
    StreamState *m_state;
    cudaStream_t m_stream;
    size_t m_base = 0;          // arena bytes in use by enclosing Scratch objects
    void *m_user_allocs[8];
    int m_user_alloc_count = 0;
    bool m_locked = true;
    uint32_t *m_pinned = nullptr;
    // allocator hooks, snapshotted under the library lock at construction
    drjit_b200_malloc_fn m_malloc_fn = nullptr;
    drjit_b200_free_fn m_free_fn = nullptr;
    void *m_alloc_user = nullptr;
};

/// Event (timing disabled) private to the calling thread and the current device
cudaEvent_t thread_event();

void count_launch();

/// Brackets one primitive call at the C-ABI boundary: launch hook (KernelHistory / LaunchBlocking,
/// src/cuda_ts.cpp:19-46) before the first and after the last launch of the call.
class CallScope {
public:
    CallScope(int kernel_type, uint32_t size, cudaStream_t stream);
    ~CallScope();
private:
    int m_type; uint32_t m_size; cudaStream_t m_stream;
    uint64_t m_launches_before;
    void *m_cookie = nullptr;
    cudaEvent_t m_start = nullptr;
    bool m_history = false, m_blocking = false, m_hooked = false;
};

/// Launch-error check + accounting after every kernel launch (submit_gpu, cuda_ts.cpp:12-47)
#define DJB_POST_LAUNCH()                                                                \
    do { DJB_CUDA_CHECK(cudaGetLastError()); ::djb::count_launch(); } while (0)

inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (uint32_t) (((uint64_t) a + b - 1) / b); }
inline uint64_t ceil_div64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
inline uint32_t round_pow2(uint32_t x) { uint32_t r = 1; while (r < x) r <<= 1; return r; }

const char *type_name(int vt);
const char *op_name(int op);
uint32_t type_size(int vt);
/// Reduction identity as raw bits (jitc_reduce_identity, src/var.cpp:2642-2652)
uint64_t reduce_identity(int vt, int op);

// ---- primitive implementations (one translation unit each) ------------------------
void memset_async(cudaStream_t s, void *ptr, uint32_t size, uint32_t isize, const void *src);
void block_reduce(cudaStream_t s, int vt, int op, uint32_t size, uint32_t block_size,
                  const void *in, void *out);
void block_reduce_bool(cudaStream_t s, const uint8_t *values, uint32_t size, uint8_t *out, int op);
bool all_any(cudaStream_t s, const uint8_t *values, uint32_t size, int op);
void reduce_dot(cudaStream_t s, int vt, const void *a, const void *b, uint32_t size, void *out);
/// Unsegmented scan of a small array in one CTA with all loads in flight (prefix_small.cu); false if the
/// configuration is left to the tile kernel
bool prefix_small(cudaStream_t s, int vt, int op, uint32_t size, bool exclusive, bool reverse, const void *in,
                  void *out, const void *carry_in, void *total_out);
/// Segmented scan of medium-sized blocks by a group of lanes per block (prefix_group.cu); false if the
/// configuration is left to the general kernel
bool prefix_group_blocks(cudaStream_t s, int vt, int op, uint32_t size, uint32_t block_size, bool exclusive,
                         bool reverse, const void *in, void *out);
void block_prefix_reduce(cudaStream_t s, int vt, int op, uint32_t size, uint32_t block_size,
                         bool exclusive, bool reverse, const void *in, void *out,
                         const void *carry_in, void *total_out);
uint32_t compress(cudaStream_t s, const uint8_t *in, uint32_t size, uint32_t index_base,
                  uint32_t *out, uint32_t *count_dev, bool sync);
uint32_t block_mkperm(cudaStream_t s, const uint32_t *values, uint32_t size, uint32_t block_size,
                      uint32_t bucket_count, uint32_t *perm, uint32_t *offsets);
uint32_t call_reduce(cudaStream_t s, const uint32_t *ids, uint32_t size, uint32_t bucket_count, uint32_t *perm,
                     uint32_t *offsets, uint32_t n_payloads, const void *const *pay_in, void *const *pay_out);
void mkperm_sharded(cudaStream_t s, const uint32_t *values, uint32_t size, uint32_t bucket_count,
                    uint32_t index_base, uint32_t *perm, uint32_t *hist_dev);
void sort(cudaStream_t s, int vt, uint32_t size, bool descending, const void *keys, void *keys_out,
          uint32_t *index_out);
void poke(cudaStream_t s, void *dst, const void *src, uint32_t size);
void aggregate(cudaStream_t s, void *dst, const drjit_b200_aggregation_entry *agg, uint32_t size);
void scatter_reduce(cudaStream_t s, int vt, int op, int mode, void *target, uint32_t target_size,
                    const void *value, const uint32_t *index, const uint8_t *mask, uint32_t size);
/// Packet form of scatter_reduce and dr.scatter_inc (scatter_packet.cu)
void scatter_reduce_packet(cudaStream_t s, int vt, int op, int mode, void *target, uint32_t target_packets,
                           const void *const *values, uint32_t count, const uint32_t *index,
                           const uint8_t *mask, uint32_t size);
void scatter_inc(cudaStream_t s, uint32_t *target, uint32_t target_size, const uint32_t *index,
                 const uint8_t *mask, uint32_t size, uint32_t *out);
void fill_fmix32(cudaStream_t s, int kind, void *out, uint64_t start, uint64_t n, uint32_t xor_,
                 uint32_t and_);

// ---- multi-GPU forms: the primitive fused with its combine over the ranks of a peer-memory
//      communicator (comm.cu / comm.cuh) ------------------------------------
struct Comm;
Comm *comm_create(uint32_t rank, uint32_t world, size_t bulk_bytes);
void comm_handle(const Comm *c, void *handle_out);
void comm_connect_ipc(Comm *c, const void *handles);
void comm_connect_local(Comm **comms, uint32_t world);
void comm_destroy(Comm *c);
uint32_t comm_rank(const Comm *c);
uint32_t comm_world(const Comm *c);
void comm_fold_scalar(cudaStream_t s, const Comm *c, int vt, int op, uint32_t fold, const void *src, void *dst);
void comm_allgather(cudaStream_t s, const Comm *c, const void *src, uint32_t bytes, void *dst);
void comm_allreduce(cudaStream_t s, const Comm *c, int vt, int op, void *data, uint32_t n);
void comm_reduce(cudaStream_t s, const Comm *c, int vt, int op, uint32_t fold, uint32_t size, const void *in, void *out);
void comm_reduce_dot(cudaStream_t s, const Comm *c, int vt, const void *a, const void *b, uint32_t size, void *out);
bool comm_all_any(cudaStream_t s, const Comm *c, const uint8_t *values, uint32_t size, int op);
void comm_prefix_reduce(cudaStream_t s, const Comm *c, int vt, int op, uint32_t size, bool exclusive, bool reverse,
                        const void *in, void *out, void *offset_out, bool materialise);
void comm_compress(cudaStream_t s, const Comm *c, const uint8_t *in, uint32_t size, uint32_t index_base,
                   uint32_t *out, uint32_t *counts_host);
uint32_t comm_mkperm(cudaStream_t s, const Comm *c, const uint32_t *values, uint32_t size, uint32_t bucket_count,
                     uint32_t index_base, uint32_t *perm, uint32_t *hist_dev, uint32_t *rank_base_dev,
                     uint32_t *offsets);

} // namespace djb
