// internal.h -- shared host-side definitions of libpasture_b200 (not part of the ABI)
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
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
    int64_t no_grouped_copy = 0;  // differential tests: columnar -> packed-record copies with one lane per record
    int64_t sort_force_8bit = 0;  // experiments: keep 8-bit digits even where 9-bit ones save a pass
    int64_t stage_chunk_mb = 0;  // staged bytes per chunk of the HOST-memspace pipeline (0 = 128 MB)
    int64_t knn_init_radius = -1, knn_stats = 0, knn_per_axis_codes = 0, knn_heap = 1;  // experiments / diagnostics
    // scratch
    void* d_scratch = nullptr;  // small device scratch (counters, partials)
    size_t d_scratch_bytes = 0;
    void* h_scratch = nullptr;  // pinned host scratch for small readbacks
    void* h_stage = nullptr;    // pinned host staging for small uploads (voxel marker tables)
    size_t h_stage_bytes = 0;
    // staging for HOST memspace buffers
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    // phase timer ("profile.phases" = 1): CUDA event pairs around the phases of a call, read by pb200_ctx_profile_read
    int64_t profile = 0;
    struct PhaseRec { const char* name; cudaEvent_t e0, e1; double host_t0_us, host_t1_us; };
    std::vector<PhaseRec> phases;
    // device-memory cache (core.cu cache_alloc / cache_free): temporaries and library-owned results of repeated calls are
    // recycled here instead of going through the driver
    std::vector<std::pair<void*, size_t>> cache_free_blocks;
    std::unordered_map<void*, size_t> cache_live;
    size_t cache_reserved = 0;
    // schedule tuning of the tile pipeline ("convert.autotune" = 1): cost-model weights that gave the fastest schedule, per plan
    // signature (convert.cu: tune_schedule)
    int64_t autotune = 0;
    struct CostModel { int64_t div = 24, pack_base = 24, pack_per_src = 12, copy_base = 6, store = 1, group_base = 4, group_store = 4, hist = 16, item = 260, load_first = -1, warp0 = 250, track = 4, cut_rows = 2; };
    std::unordered_map<uint64_t, CostModel> tuned;
    cudaEvent_t tune_e0 = nullptr, tune_e1 = nullptr;
    bool convert_attr_set = false;  // cudaFuncSetAttribute is per device: every context configures its kernels once
    bool sort_attr_set[2] = {false, false}, knn_attr_set = false, voxel_attr_set = false;
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
// Device memory for temporaries and library-owned results.  cudaMallocAsync from the driver's stream-ordered pool cost
// 1.2-1.4 ms per GB on B200 even when the pool already held the memory (measured: 2.3 ms for the two 0.8 GB key arrays
// of a 100 M-point voxel grid, 2.8 ms for its 1.2 GB of results; the pool re-maps on every request), i.e. 5 ms of a 13 ms
// call.  The context therefore keeps freed blocks itself: best fit among them (at most 25 % + 2 MiB larger than the request),
// cudaMalloc only when nothing fits.  Safe without synchronisation because all work of a context is ordered on ONE
// stream (a block is only handed out again to work queued behind its last user; pb200_ctx_set_stream orders the new
// stream behind the old one).  pb200_ctx_trim gives everything back.
cudaError_t cache_alloc(pb200_ctx* ctx, void** out, size_t bytes);
void cache_free(pb200_ctx* ctx, void* p);
void cache_trim(pb200_ctx* ctx);
extern thread_local pb200_ctx* tl_ctx;  // context of the API call running on this thread (set by DeviceGuard)
// Every entry point makes its context's device current for the duration of the call and puts the caller's device back
// on return (a host framework such as torch keeps its own notion of the current device).
struct DeviceGuard {
    int prev = -1, rc = 0;
    pb200_ctx* prev_ctx = nullptr;
    explicit DeviceGuard(pb200_ctx* ctx) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        rc = ensure_device(ctx);
        if (rc == 0 && ctx && prev == ctx->device) prev = -1;  // nothing to restore
        prev_ctx = tl_ctx;
        if (ctx) tl_ctx = ctx;
    }
    ~DeviceGuard() {
        tl_ctx = prev_ctx;
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define PB_DEVICE(ctx)                         \
    pb200::DeviceGuard _pb_device_guard(ctx);  \
    if (_pb_device_guard.rc < 0) return _pb_device_guard.rc
// device scratch of at least `bytes` (grown lazily); contents undefined
int scratch(pb200_ctx* ctx, size_t bytes, void** out);
int validate_desc(const pb200_buffer_desc* d, const char* what);

// Temporary device memory of one API call, taken from the context's cache (cache_alloc) and handed back when it goes out
// of scope: no host synchronisation, the work already queued on the context's stream that uses it runs before any later
// user of the block.  Outside an API call (no current context) it falls back to the driver's stream-ordered pool.
struct DevTmp {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    pb200_ctx* owner = nullptr;
    DevTmp() = default;
    DevTmp(const DevTmp&) = delete;
    DevTmp& operator=(const DevTmp&) = delete;
    ~DevTmp() { release(); }
    void release() {
        if (p) { if (owner) cache_free(owner, p); else cudaFreeAsync(p, st); }
        p = nullptr;
    }
    cudaError_t alloc(cudaStream_t s, size_t bytes) {
        release();
        st = s;
        owner = tl_ctx;
        if (owner) return cache_alloc(owner, &p, bytes ? bytes : 1);
        return cudaMallocAsync(&p, bytes ? bytes : 1, s);
    }
};

// Scoped phase timer: PB_PHASE(ctx, "voxel.reduce") brackets the kernels launched in the enclosing scope with two events
// on the context's stream when profiling is on (no cost otherwise) and notes the host time of scope entry / exit.
// bench.py turns the records into per-kernel milliseconds and roofline fractions of the C3 / C4 configurations.
struct PhaseScope {
    pb200_ctx* c;
    int idx = -1;
    PhaseScope(pb200_ctx* ctx, const char* name) : c(ctx) {
        if (!c || !c->profile) return;
        pb200_ctx::PhaseRec r{name, nullptr, nullptr, now_us(), 0.0};
        if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) { cudaGetLastError(); return; }
        cudaEventRecord(r.e0, c->stream);
        idx = (int)c->phases.size();
        c->phases.push_back(r);
    }
    ~PhaseScope() {
        if (idx < 0) return;
        cudaEventRecord(c->phases[(size_t)idx].e1, c->stream);
        c->phases[(size_t)idx].host_t1_us = now_us();
    }
    static double now_us() {
        return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
};
#define PB_PHASE_CAT2(a, b) a##b
#define PB_PHASE_CAT(a, b) PB_PHASE_CAT2(a, b)
#define PB_PHASE(ctx, name) pb200::PhaseScope PB_PHASE_CAT(_phase_, __LINE__)(ctx, name)

// radix_sort.cu: stable LSD radix sort by the key bits [begin_bit, end_bit); keys/vals are clobbered, the result is in
// (keys, vals) or (keys_alt, vals_alt) as *in_alt says; vals may be null (keys only); n < 2^30
int radix_sort_u64(pb200_ctx* ctx, unsigned long long* keys, unsigned long long* keys_alt, uint32_t* vals, uint32_t* vals_alt,
                   uint64_t n, int begin_bit, int end_bit, bool* in_alt);

// convert.cu: conversion into a fresh target range fused with the LAS writer's running statistics (see the definition)
struct EgressStats {
    uint64_t out_of_range;
    uint64_t hist[16];
    int bounds_tracked, has_bounds;
    double src_min[3], src_max[3];
};
int convert_range_egress(pb200_converter* cv, const pb200_buffer_desc* src, uint64_t sb, uint64_t se, const pb200_buffer_desc* dst,
                         uint64_t db, uint64_t de, int track_src_attr, int hist_src_attr, EgressStats* out);

// exclusive prefix sum of n device counters in place (one CTA; meant for per-tile counts), total -> *total_out (device)
int exclusive_scan_u32(pb200_ctx* ctx, uint32_t* counts, uint32_t n, uint32_t* total_out);

}  // namespace pb200
