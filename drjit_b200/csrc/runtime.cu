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

struct Scratch::StreamState {
    std::mutex mutex;
    void *arena = nullptr;
    size_t arena_size = 0;
    uint32_t *counters = nullptr;
    uint32_t *pinned = nullptr;
};

static std::map<std::pair<int, cudaStream_t>, Scratch::StreamState *> *g_streams = nullptr;

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
    }
    m_state->mutex.lock();
    if (!m_state->counters) {
        // A constructor that throws never runs its destructor: release the stream's lock and
        // leave the control block unallocated so that the next call starts over.
        try {
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
}

Scratch::~Scratch() {
    for (int i = 0; i < m_user_alloc_count; ++i)
        g_free_fn(m_user_allocs[i], g_alloc_user);
    m_state->mutex.unlock();
}

void Scratch::reserve(size_t total_bytes) {
    if (g_malloc_fn || m_used != 0 || total_bytes <= m_state->arena_size)
        return;
    void *p = device(total_bytes); // grows the arena
    (void) p;
    m_used = 0;
}

void *Scratch::device(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t) 255;
    if (bytes == 0)
        bytes = 256;

    if (g_malloc_fn) {
        if (m_user_alloc_count == 8)
            raise(DRJIT_B200_EFATAL, "drjit_b200: internal error (too many scratch allocations)");
        void *p = g_malloc_fn(bytes, 0, g_alloc_user);
        if (!p)
            raise(DRJIT_B200_ECUDA, "drjit_b200: scratch allocator hook returned NULL (%zu bytes)", bytes);
        m_user_allocs[m_user_alloc_count++] = p;
        return p;
    }

    if (m_used + bytes > m_state->arena_size) {
        if (m_used != 0)
            raise(DRJIT_B200_EFATAL, "drjit_b200: internal error (scratch arena grew mid-call)");
        // Grow geometrically. cudaFree() synchronises the device, so work that still
        // uses the old arena has finished before it disappears. Rare (log2 many times).
        size_t new_size = m_state->arena_size ? m_state->arena_size : (size_t) 1 << 20;
        while (new_size < bytes)
            new_size *= 2;
        if (m_state->arena)
            DJB_CUDA_CHECK(cudaFree(m_state->arena));
        m_state->arena = nullptr;
        m_state->arena_size = 0;
        DJB_CUDA_CHECK(cudaMalloc(&m_state->arena, new_size));
        m_state->arena_size = new_size;
    }
    void *p = (uint8_t *) m_state->arena + m_used;
    m_used += bytes;
    return p;
}

uint32_t *Scratch::zeroed_counters() { return m_state->counters; }
uint32_t *Scratch::pinned_words() { return m_state->pinned; }

static thread_local uint64_t t_launch_count = 0;
void count_launch() { ++t_launch_count; }

} // namespace djb

// ---------------------------------------------------------------------------
//  C ABI: library management
// ---------------------------------------------------------------------------
using namespace djb;

thread_local char t_last_error[1024] = "";

extern "C" {

DRJIT_B200_API const char *drjit_b200_last_error(void) { return t_last_error; }

DRJIT_B200_API const char *drjit_b200_version(void) { return "drjit-b200 0.1 (sm_100a)"; }

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

DRJIT_B200_API int drjit_b200_shutdown(void) {
    std::lock_guard<std::mutex> guard(g_mutex);
    if (g_streams) {
        for (auto &kv : *g_streams) {
            Scratch::StreamState *st = kv.second;
            int prev = -1;
            cudaGetDevice(&prev);
            cudaSetDevice(kv.first.first);
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
