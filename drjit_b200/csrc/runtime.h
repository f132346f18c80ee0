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
/// installed (jitc_malloc-style, released when the object dies), otherwise from a
/// grow-only arena owned by the library and keyed by (device, stream): successive
/// kernels on one stream are ordered, so the arena can be reused without a sync.
class Scratch {
public:
    Scratch(cudaStream_t stream);
    ~Scratch();
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
    /// Pinned, device-mapped host words private to this stream (results read by the host)
    uint32_t *pinned_words();
    static constexpr size_t kPinnedWords = 64;

    struct StreamState;

private:
    StreamState *m_state;
    cudaStream_t m_stream;
    size_t m_used = 0;
    void *m_user_allocs[8];
    int m_user_alloc_count = 0;
};

void count_launch();

/// Launch-error check + accounting after every kernel launch (submit_gpu, cuda_ts.cpp:12-47)
#define DJB_POST_LAUNCH()                                                                \
    do { DJB_CUDA_CHECK(cudaGetLastError()); ::djb::count_launch(); } while (0)

inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (uint32_t) (((uint64_t) a + b - 1) / b); }
inline uint64_t ceil_div64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
inline uint32_t round_pow2(uint32_t x) { uint32_t r = 1; while (r < x) r <<= 1; return r; }

const char *type_name(int vt);
const char *op_name(int op);
uint32_t type_size(int vt);

// ---- primitive implementations (one translation unit each) ------------------------
void memset_async(cudaStream_t s, void *ptr, uint32_t size, uint32_t isize, const void *src);
void block_reduce(cudaStream_t s, int vt, int op, uint32_t size, uint32_t block_size,
                  const void *in, void *out);
void block_reduce_bool(cudaStream_t s, const uint8_t *values, uint32_t size, uint8_t *out, int op);
bool all_any(cudaStream_t s, const uint8_t *values, uint32_t size, int op);
void reduce_dot(cudaStream_t s, int vt, const void *a, const void *b, uint32_t size, void *out);
void block_prefix_reduce(cudaStream_t s, int vt, int op, uint32_t size, uint32_t block_size,
                         bool exclusive, bool reverse, const void *in, void *out,
                         const void *carry_in, void *total_out);
uint32_t compress(cudaStream_t s, const uint8_t *in, uint32_t size, uint32_t index_base,
                  uint32_t *out, uint32_t *count_dev, bool sync);
uint32_t block_mkperm(cudaStream_t s, const uint32_t *values, uint32_t size, uint32_t block_size,
                      uint32_t bucket_count, uint32_t *perm, uint32_t *offsets);
void mkperm_sharded(cudaStream_t s, const uint32_t *values, uint32_t size, uint32_t bucket_count,
                    uint32_t index_base, uint32_t *perm, uint32_t *hist_dev);
void poke(cudaStream_t s, void *dst, const void *src, uint32_t size);
void aggregate(cudaStream_t s, void *dst, const drjit_b200_aggregation_entry *agg, uint32_t size);
void scatter_reduce(cudaStream_t s, int vt, int op, int mode, void *target, uint32_t target_size,
                    const void *value, const uint32_t *index, const uint8_t *mask, uint32_t size);
void fill_fmix32(cudaStream_t s, int kind, void *out, uint64_t start, uint64_t n, uint32_t xor_,
                 uint32_t and_);

} // namespace djb
