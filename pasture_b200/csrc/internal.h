// internal.h -- shared host-side definitions of libpasture_b200 (not part of the ABI)
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pasture_b200.h"

struct pb200_layout {
    std::vector<pb200_attr> attrs;
    uint64_t size = 0;   // memory_layout.size()
    uint64_t align = 1;  // memory_layout.align()
};

struct pb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int sm_count = 148;
    size_t smem_optin = 0;
    // tuning knobs (0 = automatic)
    int64_t tile_points = 0, threads = 0, stages = 0, ctas_per_sm = 0, force_direct = 0;
    int64_t stage_chunk_mb = 0;  // staged bytes per chunk of the HOST-memspace pipeline (0 = 128 MB)
    int64_t knn_init_radius = -1, knn_stats = 0, knn_per_axis_codes = 0, knn_heap = 1;  // experiments / diagnostics
    // scratch
    void* d_scratch = nullptr;  // small device scratch (counters, partials)
    size_t d_scratch_bytes = 0;
    void* h_scratch = nullptr;  // pinned host scratch for small readbacks
    // staging for HOST memspace buffers
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    // phase timer ("profile.phases" = 1): CUDA event pairs around the phases of a call, read by pb200_ctx_profile_read
    int64_t profile = 0;
    struct PhaseRec { const char* name; cudaEvent_t e0, e1; };
    std::vector<PhaseRec> phases;
    bool convert_attr_set = false;  // cudaFuncSetAttribute is per device: every context configures its kernels once
    bool sort_attr_set[2] = {false, false}, knn_attr_set = false;
};

namespace pb200 {

extern std::atomic<uint64_t> g_launches;
int set_error(int code, const char* fmt, ...);
int cuda_error(cudaError_t e, const char* what);

#define PB_CUDA(x)                                         \
    do {                                                   \
        cudaError_t _e = (x);                              \
        if (_e != cudaSuccess) return pb200::cuda_error(_e, #x); \
    } while (0)
#define PB_TRY(x)            \
    do {                     \
        int _rc = (x);       \
        if (_rc < 0) return _rc; \
    } while (0)

inline bool is_scalar(uint32_t d) { return d <= PB200_F64; }
inline bool is_cast_vec3(uint32_t d) { return d >= PB200_VEC3U8 && d <= PB200_VEC3F64; }
inline uint32_t vec3_component(uint32_t d) {
    switch (d) {
        case PB200_VEC3U8: return PB200_U8;
        case PB200_VEC3U16: return PB200_U16;
        case PB200_VEC3F32: return PB200_F32;
        case PB200_VEC3I32: return PB200_I32;
        default: return PB200_F64;
    }
}
inline bool has_conversion(uint32_t from, uint32_t to) {  // attribute_conversion.rs:194-260
    if (from == to) return false;
    return (is_scalar(from) && is_scalar(to)) || (is_cast_vec3(from) && is_cast_vec3(to));
}
inline bool dtype_equal(const pb200_attr& a, const pb200_attr& b) {
    if (a.dtype != b.dtype) return false;
    if (a.dtype == PB200_BYTEARRAY || a.dtype == PB200_CUSTOM) return a.extra_size == b.extra_size;
    return true;
}
int ensure_device(pb200_ctx* ctx);
// Every entry point makes its context's device current for the duration of the call and puts the caller's device back
// on return (a host framework such as torch keeps its own notion of the current device).
struct DeviceGuard {
    int prev = -1, rc = 0;
    explicit DeviceGuard(pb200_ctx* ctx) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        rc = ensure_device(ctx);
        if (rc == 0 && ctx && prev == ctx->device) prev = -1;  // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define PB_DEVICE(ctx)                         \
    pb200::DeviceGuard _pb_device_guard(ctx);  \
    if (_pb_device_guard.rc < 0) return _pb_device_guard.rc
// device scratch of at least `bytes` (grown lazily); contents undefined
int scratch(pb200_ctx* ctx, size_t bytes, void** out);
int validate_desc(const pb200_buffer_desc* d, const char* what);

// Stream-ordered temporary (cudaMallocAsync from the device's default pool, whose release threshold the context raises
// so that the multi-GB sort / tree buffers of repeated calls are recycled instead of going back to the driver).
// Freed behind the work already queued on the stream: no host synchronisation is needed before it goes out of scope.
struct DevTmp {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    DevTmp() = default;
    DevTmp(const DevTmp&) = delete;
    DevTmp& operator=(const DevTmp&) = delete;
    ~DevTmp() { release(); }
    void release() { if (p) cudaFreeAsync(p, st); p = nullptr; }
    cudaError_t alloc(cudaStream_t s, size_t bytes) {
        release();
        st = s;
        return cudaMallocAsync(&p, bytes ? bytes : 1, s);
    }
};

// Scoped phase timer: PB_PHASE(ctx, "voxel.reduce") brackets the kernels launched in the enclosing scope with two events
// on the context's stream when profiling is on (no cost otherwise).  bench.py turns the records into per-kernel
// milliseconds and roofline fractions of the C3 / C4 configurations.
struct PhaseScope {
    pb200_ctx* c;
    int idx = -1;
    PhaseScope(pb200_ctx* ctx, const char* name) : c(ctx) {
        if (!c || !c->profile) return;
        pb200_ctx::PhaseRec r{name, nullptr, nullptr};
        if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) { cudaGetLastError(); return; }
        cudaEventRecord(r.e0, c->stream);
        idx = (int)c->phases.size();
        c->phases.push_back(r);
    }
    ~PhaseScope() { if (idx >= 0) cudaEventRecord(c->phases[(size_t)idx].e1, c->stream); }
};
#define PB_PHASE_CAT2(a, b) a##b
#define PB_PHASE_CAT(a, b) PB_PHASE_CAT2(a, b)
#define PB_PHASE(ctx, name) pb200::PhaseScope PB_PHASE_CAT(_phase_, __LINE__)(ctx, name)

// radix_sort.cu: stable LSD radix sort by the key bits [begin_bit, end_bit); keys/vals are clobbered, the result is in
// (keys, vals) or (keys_alt, vals_alt) as *in_alt says; vals may be null (keys only); n < 2^30
int radix_sort_u64(pb200_ctx* ctx, unsigned long long* keys, unsigned long long* keys_alt, uint32_t* vals, uint32_t* vals_alt,
                   uint64_t n, int begin_bit, int end_bit, bool* in_alt);

// exclusive prefix sum of n device counters in place (one CTA; meant for per-tile counts), total -> *total_out (device)
int exclusive_scan_u32(pb200_ctx* ctx, uint32_t* counts, uint32_t n, uint32_t* total_out);

}  // namespace pb200
