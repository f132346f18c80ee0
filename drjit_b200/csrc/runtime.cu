/*
 * runtime.cu -- see runtime.h. Host-only code (compiled by nvcc for convenience).
 */
#include "runtime.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace djb {

void raise(int code, const char *fmt, ...) {
    char buf[1024];
    va_list args;
    va_start(args, fmt);
    vsnprintf(buf, sizeof(buf), fmt, args);
    va_end(args);
    throw Error(code, buf);
}

// ---------------------------------------------------------------------------
//  Names (wording follows src/util.cpp:22-23 and src/var.cpp type_name[])
// ---------------------------------------------------------------------------
const char *type_name(int vt) {
    switch (vt) {
        case DRJIT_B200_VT_BOOL: return "bool";     case DRJIT_B200_VT_INT8: return "int8";
        case DRJIT_B200_VT_UINT8: return "uint8";   case DRJIT_B200_VT_INT16: return "int16";
        case DRJIT_B200_VT_UINT16: return "uint16"; case DRJIT_B200_VT_INT32: return "int32";
        case DRJIT_B200_VT_UINT32: return "uint32"; case DRJIT_B200_VT_INT64: return "int64";
        case DRJIT_B200_VT_UINT64: return "uint64"; case DRJIT_B200_VT_FLOAT16: return "float16";
        case DRJIT_B200_VT_FLOAT32: return "float32"; case DRJIT_B200_VT_FLOAT64: return "float64";
        default: return "invalid";
    }
}

const char *op_name(int op) {
    static const char *names[] = { "identity", "add", "mul", "min", "max", "and", "or" };
    return (op >= 0 && op <= 6) ? names[op] : "invalid";
}

uint32_t type_size(int vt) {
    switch (vt) {
        case DRJIT_B200_VT_BOOL: case DRJIT_B200_VT_INT8: case DRJIT_B200_VT_UINT8: return 1;
        case DRJIT_B200_VT_INT16: case DRJIT_B200_VT_UINT16: case DRJIT_B200_VT_FLOAT16: return 2;
        case DRJIT_B200_VT_INT32: case DRJIT_B200_VT_UINT32: case DRJIT_B200_VT_FLOAT32: return 4;
        case DRJIT_B200_VT_INT64: case DRJIT_B200_VT_UINT64: case DRJIT_B200_VT_FLOAT64: return 8;
        default: return 0;
    }
}

/// Reduction identity as raw bits (jitc_reduce_identity, src/var.cpp:2642-2652)
uint64_t reduce_identity(int vt, int op) {
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

// ---------------------------------------------------------------------------
//  Device table
// ---------------------------------------------------------------------------
static std::mutex g_mutex;
static std::map<int, DeviceProps> g_devices;

const DeviceProps &device_props() {
    int device = -1;
    DJB_CUDA_CHECK(cudaGetDevice(&device));
    std::lock_guard<std::mutex> guard(g_mutex);
    auto it = g_devices.find(device);
    if (it != g_devices.end())
        return it->second;

    DeviceProps p;
    int v = 0, major = 0, minor = 0;
    p.device = device;
    DJB_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    p.sm_count = (uint32_t) v;
    DJB_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    p.smem_optin = (uint32_t) v;
    DJB_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    DJB_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    p.cc = (uint32_t) (major * 10 + minor);
    if (major != 10)
        raise(DRJIT_B200_ECUDA,
              "drjit_b200: device %i has compute capability %i.%i, but this library only "
              "contains sm_100a (B200) kernels and has no fallback path.", device, major, minor);
    return g_devices.emplace(device, p).first->second;
}

// ---------------------------------------------------------------------------
//  Allocator hooks + per-stream scratch state
// ---------------------------------------------------------------------------
static drjit_b200_malloc_fn g_malloc_fn = nullptr;
static drjit_b200_free_fn g_free_fn = nullptr;
static void *g_alloc_user = nullptr;

struct Retired { void *ptr; cudaEvent_t done; };

struct Scratch::StreamState {
    // Recursive: a primitive that is built from other primitives (the sharded scan = reduction +
    // scan) keeps its own Scratch alive across the inner calls; inner objects continue where the
    // outer one stopped (`used`) and give their part back when they die, like a stack.
    std::recursive_mutex mutex;
    uint32_t depth = 0;               // live Scratch objects of the thread that holds the mutex
    size_t used = 0;                  // bytes of the arena handed out to them
    void *arena = nullptr;
    size_t arena_size = 0;
    uint32_t *counters = nullptr;
    uint32_t *pinned = nullptr;
    uint32_t next_slot = 0;
    std::vector<Retired> retired;     // outgrown arenas, freed once the work that used them is done
};

static std::map<std::pair<int, cudaStream_t>, Scratch::StreamState *> *g_streams = nullptr;

/// Free outgrown arenas whose last user has finished (never inside a stream capture)
static void collect_retired(Scratch::StreamState *st, bool wait, bool shutdown = false) {
    for (size_t i = 0; i < st->retired.size();) {
        Retired &r = st->retired[i];
        if (!r.done) {                      // still referenced by a live call (or untrackable)
            if (!shutdown) { ++i; continue; }
            cudaFree(r.ptr);                // shutdown: streams are idle by contract
            st->retired.erase(st->retired.begin() + (long) i);
            continue;
        }
        cudaError_t rv = wait ? cudaEventSynchronize(r.done) : cudaEventQuery(r.done);
        if (rv == cudaSuccess) {
            cudaFree(r.ptr);
            cudaEventDestroy(r.done);
            st->retired.erase(st->retired.begin() + (long) i);
        } else {
            (void) cudaGetLastError();      // cudaErrorNotReady is not sticky, but clear it anyway
            ++i;
        }
    }
}

static bool stream_is_capturing(cudaStream_t stream) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &status) != cudaSuccess) {
        (void) cudaGetLastError();
        return false;
    }
    return status != cudaStreamCaptureStatusNone;
}

Scratch::Scratch(cudaStream_t stream) : m_stream(stream) {
    int device = device_props().device;
    {
        std::lock_guard<std::mutex> guard(g_mutex);
        if (!g_streams)
            g_streams = new std::map<std::pair<int, cudaStream_t>, StreamState *>();
        StreamState *&st = (*g_streams)[std::make_pair(device, stream)];
        if (!st)
            st = new StreamState();
        m_state = st;
        m_malloc_fn = g_malloc_fn; m_free_fn = g_free_fn; m_alloc_user = g_alloc_user;
    }
    m_state->mutex.lock();
    if (!m_state->counters) {
        // A constructor that throws never runs its destructor: release the stream's lock and
        // leave the control block unallocated so that the next call starts over.
        try {
            if (stream_is_capturing(stream))
                raise(DRJIT_B200_ECUDA, "drjit_b200: the first call on a stream allocates its control block and "
                                        "cannot be captured into a CUDA graph; call drjit_b200_reserve_scratch() "
                                        "(or run the primitive once) before the capture.");
            uint32_t *counters = nullptr, *pinned = nullptr;
            DJB_CUDA_CHECK(cudaMalloc((void **) &counters, kZeroedCounters * sizeof(uint32_t)));
            cudaError_t rv = cudaMemset(counters, 0, kZeroedCounters * sizeof(uint32_t));
            if (rv == cudaSuccess)
                rv = cudaHostAlloc((void **) &pinned, kPinnedWords * sizeof(uint32_t),
                                   cudaHostAllocMapped | cudaHostAllocPortable);
            if (rv != cudaSuccess) {
                cudaFree(counters);
                DJB_CUDA_CHECK(rv);
            }
            memset(pinned, 0, kPinnedWords * sizeof(uint32_t));
            m_state->counters = counters;
            m_state->pinned = pinned;
        } catch (...) {
            m_state->mutex.unlock();
            throw;
        }
    }
    m_pinned = m_state->pinned + kPinnedSlotWords * (m_state->next_slot++ % kPinnedSlots);
    m_base = m_state->used;
    ++m_state->depth;
}

void Scratch::unlock() {
    if (!m_locked)
        return;
    // Temporaries from the allocator hooks go back now: the hook's free is stream-ordered (the
    // contract in drjit_b200.h), the kernels using them were enqueued before this point.
    for (int i = 0; i < m_user_alloc_count; ++i)
        m_free_fn(m_user_allocs[i], m_alloc_user);
    m_user_alloc_count = 0;
    m_locked = false;
    m_state->used = m_base;
    if (--m_state->depth == 0) {
        // arenas outgrown during this call: their last users are enqueued by now
        for (Retired &r : m_state->retired)
            if (!r.done && (cudaEventCreateWithFlags(&r.done, cudaEventDisableTiming) != cudaSuccess ||
                            cudaEventRecord(r.done, m_stream) != cudaSuccess)) {
                (void) cudaGetLastError();      // cannot track it: keep it until shutdown
            }
    }
    m_state->mutex.unlock();
}

Scratch::~Scratch() { unlock(); }

void Scratch::reserve(size_t total_bytes) {
    if (m_malloc_fn || m_state->used + total_bytes <= m_state->arena_size)
        return;
    const size_t before = m_state->used;
    void *p = device(total_bytes); // grows the arena
    (void) p;
    m_state->used = before;
}

void *Scratch::device(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t) 255;
    if (bytes == 0)
        bytes = 256;
    if (!m_locked)
        raise(DRJIT_B200_EFATAL, "drjit_b200: internal error (scratch requested after unlock)");

    if (m_malloc_fn) {
        if (m_user_alloc_count == 8)
            raise(DRJIT_B200_EFATAL, "drjit_b200: internal error (too many scratch allocations)");
        void *p = m_malloc_fn(bytes, 0, m_alloc_user);
        if (!p)
            raise(DRJIT_B200_ECUDA, "drjit_b200: scratch allocator hook returned NULL (%zu bytes)", bytes);
        m_user_allocs[m_user_alloc_count++] = p;
        return p;
    }

    if (m_state->used + bytes > m_state->arena_size) {
        if (m_state->used != m_base)
            raise(DRJIT_B200_EFATAL, "drjit_b200: internal error (scratch arena grew mid-call; reserve() first)");
        if (stream_is_capturing(m_stream))
            raise(DRJIT_B200_ECUDA, "drjit_b200: the scratch arena of this stream holds %zu bytes but the call needs "
                                    "%zu, and it cannot grow while the stream is being captured into a CUDA graph; "
                                    "call drjit_b200_reserve_scratch() (or run the primitive once) before the capture.",
                  m_state->arena_size, bytes);
        // Grow geometrically (rare: log2 many times). The old arena may still be in use by
        // kernels in flight on this stream: retire it behind an event instead of freeing it now
        // (cudaFree would synchronise the whole device).
        // (an enclosing Scratch keeps its pointers into the old arena: it is only retired, and the
        // event that releases it is recorded when the outermost object unlocks)
        size_t new_size = m_state->arena_size ? m_state->arena_size : (size_t) 1 << 20;
        while (new_size < m_state->used + bytes)
            new_size *= 2;
        collect_retired(m_state, false);
        void *fresh = nullptr;
        cudaError_t rv = cudaMalloc(&fresh, new_size);
        if (rv != cudaSuccess) {            // out of memory: release what is retired, then retry once
            (void) cudaGetLastError();
            collect_retired(m_state, true);
            DJB_CUDA_CHECK(cudaMalloc(&fresh, new_size));
        }
        if (m_state->arena)
            m_state->retired.push_back(Retired{ m_state->arena, nullptr });   // event: see unlock()
        m_state->arena = fresh;
        m_state->arena_size = new_size;
    }
    void *p = (uint8_t *) m_state->arena + m_state->used;
    m_state->used += bytes;
    return p;
}

uint32_t *Scratch::zeroed_counters() { return m_state->counters; }
uint32_t *Scratch::pinned_words() { return m_pinned; }

cudaEvent_t thread_event() {
    static thread_local cudaEvent_t events[kMaxDevices] = {};
    const int device = device_props().device % kMaxDevices;
    if (!events[device])
        DJB_CUDA_CHECK(cudaEventCreateWithFlags(&events[device], cudaEventDisableTiming));
    return events[device];
}

// ---------------------------------------------------------------------------
//  Launch accounting, kernel history, launch blocking (src/cuda_ts.cpp:12-47)
// ---------------------------------------------------------------------------
static thread_local uint64_t t_launch_count = 0;     // since the last reset (C ABI)
static thread_local uint64_t t_launch_total = 0;     // monotonic (CallScope)
void count_launch() { ++t_launch_count; ++t_launch_total; }

static drjit_b200_launch_hook g_launch_hook = nullptr;
static void *g_launch_hook_user = nullptr;
static thread_local uint32_t t_flags = 0;

struct HistoryRecord { uint32_t type, size, launches; cudaEvent_t start, end; };
static thread_local std::vector<HistoryRecord> *t_history = nullptr;
static thread_local int t_scope_depth = 0;           // primitives that call other primitives: outermost wins

CallScope::CallScope(int kernel_type, uint32_t size, cudaStream_t stream)
    : m_type(kernel_type), m_size(size), m_stream(stream), m_launches_before(t_launch_total) {
    if (t_scope_depth++ != 0)
        return;
    m_history = (t_flags & DRJIT_B200_FLAG_KERNEL_HISTORY) != 0;
    m_blocking = (t_flags & DRJIT_B200_FLAG_LAUNCH_BLOCKING) != 0;
    drjit_b200_launch_hook hook;
    void *user;
    {
        std::lock_guard<std::mutex> guard(g_mutex);
        hook = g_launch_hook; user = g_launch_hook_user;
    }
    if (hook) {
        m_hooked = true;
        hook(user, 0, m_type, m_size, (void *) m_stream, 0, &m_cookie);
    }
    if (m_history && !stream_is_capturing(stream)) {
        if (cudaEventCreate(&m_start) != cudaSuccess || cudaEventRecord(m_start, stream) != cudaSuccess) {
            (void) cudaGetLastError();
            if (m_start) cudaEventDestroy(m_start);
            m_start = nullptr;
        }
    }
}

CallScope::~CallScope() {
    if (--t_scope_depth != 0)
        return;
    const uint32_t launches = (uint32_t) (t_launch_total - m_launches_before);
    if (m_blocking && launches && !stream_is_capturing(m_stream))
        (void) cudaStreamSynchronize(m_stream);                     // cuda_ts.cpp:33-34
    if (m_start) {
        cudaEvent_t end = nullptr;
        if (launches && cudaEventCreate(&end) == cudaSuccess && cudaEventRecord(end, m_stream) == cudaSuccess) {
            if (!t_history)
                t_history = new std::vector<HistoryRecord>();
            t_history->push_back(HistoryRecord{ (uint32_t) m_type, m_size, launches, m_start, end });
        } else {
            (void) cudaGetLastError();
            if (end) cudaEventDestroy(end);
            cudaEventDestroy(m_start);
        }
    }
    if (m_hooked) {
        drjit_b200_launch_hook hook;
        void *user;
        {
            std::lock_guard<std::mutex> guard(g_mutex);
            hook = g_launch_hook; user = g_launch_hook_user;
        }
        if (hook)
            hook(user, 1, m_type, m_size, (void *) m_stream, launches, &m_cookie);
    }
}

} // namespace djb

// ---------------------------------------------------------------------------
//  C ABI: library management
// ---------------------------------------------------------------------------
using namespace djb;

thread_local char t_last_error[1024] = "";

extern "C" {

DRJIT_B200_API const char *drjit_b200_last_error(void) { return t_last_error; }

DRJIT_B200_API const char *drjit_b200_version(void) { return "drjit-b200 0.2 (sm_100a)"; }

DRJIT_B200_API uint64_t drjit_b200_launch_count(int reset) {
    uint64_t v = t_launch_count;
    if (reset)
        t_launch_count = 0;
    return v;
}

DRJIT_B200_API int drjit_b200_set_allocator(drjit_b200_malloc_fn malloc_fn,
                                            drjit_b200_free_fn free_fn, void *user) {
    if ((malloc_fn == nullptr) != (free_fn == nullptr)) {
        snprintf(t_last_error, sizeof(t_last_error),
                 "drjit_b200_set_allocator(): malloc and free hooks must be set together");
        return DRJIT_B200_EINVAL;
    }
    std::lock_guard<std::mutex> guard(g_mutex);
    g_malloc_fn = malloc_fn;
    g_free_fn = free_fn;
    g_alloc_user = user;
    return DRJIT_B200_OK;
}

DRJIT_B200_API int drjit_b200_set_launch_hook(drjit_b200_launch_hook hook, void *user) {
    std::lock_guard<std::mutex> guard(g_mutex);
    g_launch_hook = hook;
    g_launch_hook_user = user;
    return DRJIT_B200_OK;
}

DRJIT_B200_API int drjit_b200_set_flags(uint32_t flags) { t_flags = flags; return DRJIT_B200_OK; }
DRJIT_B200_API uint32_t drjit_b200_flags(void) { return t_flags; }

DRJIT_B200_API uint32_t drjit_b200_kernel_history(struct drjit_b200_history_entry *entries, uint32_t max_entries) {
    if (!t_history)
        return 0;
    uint32_t n = 0;
    for (HistoryRecord &r : *t_history) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.end) != cudaSuccess || cudaEventElapsedTime(&ms, r.start, r.end) != cudaSuccess) {
            (void) cudaGetLastError();
            ms = 0.f;
        }
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.end);
        if (entries && n < max_entries)
            entries[n++] = drjit_b200_history_entry{ r.type, r.size, r.launches, ms };
    }
    t_history->clear();
    return n;
}

DRJIT_B200_API void drjit_b200_kernel_history_clear(void) { (void) drjit_b200_kernel_history(nullptr, 0); }

DRJIT_B200_API int drjit_b200_reserve_scratch(void *stream, size_t bytes) {
    try {
        Scratch scratch((cudaStream_t) stream);
        scratch.reserve(bytes);
        t_last_error[0] = '\0';
        return DRJIT_B200_OK;
    } catch (const djb::Error &e) {
        snprintf(t_last_error, sizeof(t_last_error), "%s", e.what());
        return e.code;
    }
}

DRJIT_B200_API int drjit_b200_shutdown(void) {
    std::lock_guard<std::mutex> guard(g_mutex);
    if (g_streams) {
        for (auto &kv : *g_streams) {
            Scratch::StreamState *st = kv.second;
            int prev = -1;
            cudaGetDevice(&prev);
            cudaSetDevice(kv.first.first);
            collect_retired(st, true, true);
            if (st->arena) cudaFree(st->arena);
            if (st->counters) cudaFree(st->counters);
            if (st->pinned) cudaFreeHost(st->pinned);
            if (prev >= 0) cudaSetDevice(prev);
            delete st;
        }
        delete g_streams;
        g_streams = nullptr;
    }
    g_devices.clear();
    return DRJIT_B200_OK;
}

} // extern "C"
