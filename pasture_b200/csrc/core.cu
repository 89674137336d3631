// core.cu -- errors, context, memory helpers and the PointLayout descriptor (host side)
#include "internal.h"

#include <sched.h>

#include <cctype>

namespace pb200 {
extern pb200_ctx::CostModel g_cost;  // convert.cu cost model

std::atomic<uint64_t> g_launches{0};
thread_local pb200_ctx* tl_ctx = nullptr;

cudaError_t cache_alloc(pb200_ctx* ctx, void** out, size_t bytes) {
    *out = nullptr;
    // size classes: multiples of 2 MiB from 2 MiB up, powers of two (>= 512 B) below
    size_t need = bytes ? bytes : 1;
    if (need >= ((size_t)2 << 20)) need = (need + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    else { size_t c = 512; while (c < need) c <<= 1; need = c; }
    auto& fb = ctx->cache_free_blocks;
    size_t best = fb.size();
    const size_t limit = need + need / 4 + ((size_t)2 << 20);
    for (size_t i = 0; i < fb.size(); ++i)
        if (fb[i].second >= need && fb[i].second <= limit && (best == fb.size() || fb[i].second < fb[best].second)) best = i;
    if (best != fb.size()) {
        *out = fb[best].first;
        ctx->cache_live[*out] = fb[best].second;
        fb[best] = fb.back();
        fb.pop_back();
        return cudaSuccess;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaErrorMemoryAllocation) {  // give the idle blocks back and try once more
        cudaGetLastError();
        cudaStreamSynchronize(ctx->stream);
        for (auto& b : fb) { cudaFree(b.first); ctx->cache_reserved -= b.second; }
        fb.clear();
        e = cudaMalloc(&p, need);
    }
    if (e != cudaSuccess) return e;
    ctx->cache_reserved += need;
    ctx->cache_live[p] = need;
    *out = p;
    return cudaSuccess;
}

void cache_free(pb200_ctx* ctx, void* p) {
    if (!p) return;
    auto it = ctx->cache_live.find(p);
    if (it == ctx->cache_live.end()) { cudaFreeAsync(p, ctx->stream); return; }  // not ours: driver pool memory
    ctx->cache_free_blocks.emplace_back(p, it->second);
    ctx->cache_live.erase(it);
}

void cache_trim(pb200_ctx* ctx) {
    for (auto& b : ctx->cache_free_blocks) { cudaFree(b.first); ctx->cache_reserved -= b.second; }
    ctx->cache_free_blocks.clear();
}
static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_error(cudaError_t e, const char* what) {
    int code = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
                   ? PB200_ERR_NO_DEVICE
                   : (e == cudaErrorMemoryAllocation ? PB200_ERR_OOM : PB200_ERR_CUDA);
    cudaGetLastError();
    return set_error(code, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

int ensure_device(pb200_ctx* ctx) {
    if (!ctx) return set_error(PB200_ERR_INVALID, "null context");
    PB_CUDA(cudaSetDevice(ctx->device));
    return PB200_OK;
}

int scratch(pb200_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->d_scratch_bytes) {
        if (ctx->d_scratch) {
            PB_CUDA(cudaStreamSynchronize(ctx->stream));
            PB_CUDA(cudaFree(ctx->d_scratch));
            ctx->d_scratch = nullptr;
            ctx->d_scratch_bytes = 0;
        }
        size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
        PB_CUDA(cudaMalloc(&ctx->d_scratch, want));
        ctx->d_scratch_bytes = want;
    }
    *out = ctx->d_scratch;
    return PB200_OK;
}

int validate_desc(const pb200_buffer_desc* d, const char* what) {
    if (!d || !d->layout) return set_error(PB200_ERR_INVALID, "%s: null buffer descriptor/layout", what);
    if (d->kind != PB200_INTERLEAVED && d->kind != PB200_COLUMNAR)
        return set_error(PB200_ERR_INVALID, "%s: buffer must be interleaved or columnar", what);
    if (d->memspace != PB200_HOST && d->memspace != PB200_DEVICE)
        return set_error(PB200_ERR_INVALID, "%s: bad memspace", what);
    if (d->len > 0) {
        if (d->kind == PB200_INTERLEAVED && !d->aos && d->layout->size > 0)
            return set_error(PB200_ERR_INVALID, "%s: interleaved buffer without memory", what);
        if (d->kind == PB200_COLUMNAR && !d->columns && !d->layout->attrs.empty())
            return set_error(PB200_ERR_INVALID, "%s: columnar buffer without columns", what);
    }
    return PB200_OK;
}

static uint64_t align_to(uint64_t v, uint64_t a) {  // math/arithmetic.rs Alignable::align_to
    if (a == 0) return v;
    uint64_t r = v % a;
    return r == 0 ? v : v + (a - r);
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_abi_version(void) { return PB200_ABI_VERSION; }
const char* pb200_last_error(void) { return g_err; }
uint64_t pb200_kernel_launch_count(void) { return g_launches.load(); }

int pb200_ctx_create(int device, pb200_ctx** out) {
    if (!out) return set_error(PB200_ERR_INVALID, "pb200_ctx_create: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error(PB200_ERR_NO_DEVICE,
                         "no CUDA device available (%s); pasture_b200 has no CPU fallback",
                         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return set_error(PB200_ERR_NO_DEVICE, "device %d out of range (%d devices)", device, n);
    struct Restore { int prev = -1; ~Restore() { if (prev >= 0) cudaSetDevice(prev); } } restore;
    if (cudaGetDevice(&restore.prev) != cudaSuccess) { restore.prev = -1; cudaGetLastError(); }
    PB_CUDA(cudaSetDevice(device));
    pb200_ctx* c = new pb200_ctx();
    c->device = device;
    cudaDeviceProp prop;
    PB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        delete c;
        return set_error(PB200_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                         prop.major, prop.minor);
    }
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    PB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->owns_stream = true;
    PB_CUDA(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
    PB_CUDA(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
    PB_CUDA(cudaHostAlloc(&c->h_scratch, 4096, cudaHostAllocDefault));
    c->h_stage_bytes = (size_t)1 << 20;
    PB_CUDA(cudaHostAlloc(&c->h_stage, c->h_stage_bytes, cudaHostAllocDefault));
    {   // keep freed temporaries (DevTmp) in the pool between calls; pb200_ctx_trim hands them back
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    *out = c;
    return PB200_OK;
}

int pb200_ctx_set_stream(pb200_ctx* ctx, void* s) {
    PB_DEVICE(ctx);
    if ((cudaStream_t)s == ctx->stream) return PB200_OK;
    if (ctx->owns_stream && ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    } else if (ctx->stream || !ctx->cache_free_blocks.empty()) {
        // cached blocks may still be in use by work queued on the old stream: the new stream starts behind it
        cudaEvent_t ev;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
            cudaEventRecord(ev, ctx->stream);
            cudaStreamWaitEvent((cudaStream_t)s, ev, 0);
            cudaEventDestroy(ev);
        }
        cudaGetLastError();
    }
    ctx->stream = (cudaStream_t)s;
    ctx->owns_stream = false;
    return PB200_OK;
}

void* pb200_ctx_get_stream(pb200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int pb200_ctx_synchronize(pb200_ctx* ctx) {
    PB_DEVICE(ctx);
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
}

int pb200_ctx_trim(pb200_ctx* ctx) {
    PB_DEVICE(ctx);
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    cache_trim(ctx);
    cudaMemPool_t pool;
    PB_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
    PB_CUDA(cudaMemPoolTrimTo(pool, 0));
    return PB200_OK;
}

int pb200_ctx_set_param(pb200_ctx* ctx, const char* key, int64_t v) {
    if (!ctx || !key) return set_error(PB200_ERR_INVALID, "pb200_ctx_set_param: null argument");
    std::string k(key);
    if (k == "convert.tile_points") ctx->tile_points = v;
    else if (k == "convert.threads") ctx->threads = v;
    else if (k == "convert.stages") ctx->stages = v;
    else if (k == "convert.ctas_per_sm") ctx->ctas_per_sm = v;
    else if (k == "convert.force_direct") ctx->force_direct = v;
    else if (k == "convert.no_grouped_copy") ctx->no_grouped_copy = v;
    else if (k == "convert.autotune") { ctx->autotune = v; if (!v) ctx->tuned.clear(); }
    else if (k == "convert.stage_chunk_mb") ctx->stage_chunk_mb = v;
    else if (k == "convert.cost_div") g_cost.div = v;
    else if (k == "convert.cost_pack_base") g_cost.pack_base = v;
    else if (k == "convert.cost_pack_per_src") g_cost.pack_per_src = v;
    else if (k == "convert.cost_copy_base") g_cost.copy_base = v;
    else if (k == "convert.cost_store") g_cost.store = v;
    else if (k == "convert.cost_group_base") g_cost.group_base = v;
    else if (k == "convert.cost_group_store") g_cost.group_store = v;
    else if (k == "convert.cost_hist") g_cost.hist = v;
    else if (k == "convert.cost_item") g_cost.item = v;
    else if (k == "convert.load_first") g_cost.load_first = v;
    else if (k == "convert.cost_warp0") g_cost.warp0 = v;
    else if (k == "convert.cost_track") g_cost.track = v;
    else if (k == "convert.cut_rows") g_cost.cut_rows = v > 0 ? v : 1;
    else if (k == "profile.phases") ctx->profile = v;
    else if (k == "sort.force_8bit") ctx->sort_force_8bit = v;
    else if (k == "knn.init_radius") ctx->knn_init_radius = v;
    else if (k == "knn.stats") {
#ifdef PB200_KNN_DIAGNOSTICS
        ctx->knn_stats = v;
#else
        if (v) return set_error(PB200_ERR_UNSUPPORTED, "knn.stats needs a diagnostics build (make -C pasture_b200/csrc KNN_DIAGNOSTICS=1): "
                                "the release library never writes traversal counters into pb200_knn's idx_out");
#endif
    }
    else if (k == "knn.heap") ctx->knn_heap = v;
    else if (k == "knn.per_axis_codes") ctx->knn_per_axis_codes = v;
    else return set_error(PB200_ERR_INVALID, "unknown parameter %s", key);
    return PB200_OK;
}

int pb200_ctx_profile_read(pb200_ctx* ctx, char* out, uint64_t capacity) {
    PB_DEVICE(ctx);
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t used = 0;
    int n = 0;
    if (out && capacity) out[0] = 0;
    for (auto& r : ctx->phases) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess && out && used + 96 < capacity) {
            // name, device milliseconds between the two events, host time of scope entry / exit relative to the first record
            used += (uint64_t)snprintf(out + used, (size_t)(capacity - used), "%s\t%.6f\t%.1f\t%.1f\n", r.name, (double)ms,
                                       r.host_t0_us - ctx->phases[0].host_t0_us, r.host_t1_us - ctx->phases[0].host_t0_us);
            ++n;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    {   // where the temporaries live: the device's stream-ordered pool
        cudaMemPool_t pool;
        uint64_t reserved = 0, used = 0, high = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess && out && used + 160 < capacity) {
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemHigh, &high);
            size_t len = strlen(out);
            snprintf(out + len, (size_t)(capacity - len), "pool.driver_reserved_MB\t%.1f\t0\t0\npool.cache_reserved_MB\t%.1f\t0\t0\npool.cache_idle_blocks\t%.1f\t0\t0\n",
                     reserved / 1048576.0, ctx->cache_reserved / 1048576.0, (double)ctx->cache_free_blocks.size());
        }
    }
    cudaGetLastError();
    ctx->phases.clear();
    return n;
}

void pb200_ctx_destroy(pb200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& r : ctx->phases) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    cache_trim(ctx);
    for (auto& kv : ctx->cache_live) cudaFree(kv.first);  // blocks whose owners outlived the context
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    delete ctx;
}

// Host memory that a GPU reads or writes over PCIe should live on the NUMA node the GPU hangs off: with several GPUs per
// box, pinned buffers that all sit on one socket (wherever the launcher happened to start the processes) make every
// copy of the far GPUs cross the inter-socket link.  Linux places pages on the node of the thread that first touches
// them (cudaHostAlloc touches at allocation), so binding the calling thread to the device's local CPUs BEFORE it
// allocates its pinned buffers is enough; no libnuma needed.  sysfs: /sys/bus/pci/devices/<bus id>/{local_cpulist,numa_node}.
int pb200_ctx_bind_host_thread(pb200_ctx* ctx, int* numa_node_out, int* n_cpus_out) {
    PB_DEVICE(ctx);
    if (numa_node_out) *numa_node_out = -1;
    if (n_cpus_out) *n_cpus_out = 0;
    char bus[32] = "";
    PB_CUDA(cudaDeviceGetPCIBusId(bus, (int)sizeof bus, ctx->device));
    for (char* c = bus; *c; ++c) *c = (char)tolower((unsigned char)*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    if (FILE* f = fopen(path, "r")) {
        int node = -1;
        if (fscanf(f, "%d", &node) == 1 && numa_node_out) *numa_node_out = node;
        fclose(f);
    }
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = fopen(path, "r");
    if (!f) return PB200_OK;  // no sysfs (or a VM that hides the topology): leave the thread where it is
    char list[4096] = "";
    const bool got = fgets(list, sizeof list, f) != nullptr;
    fclose(f);
    if (!got) return PB200_OK;
    cpu_set_t allowed, want;
    CPU_ZERO(&allowed);
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return PB200_OK;
    int n = 0;
    for (const char* p = list; *p;) {  // "0-31,64-95"
        while (*p && !isdigit((unsigned char)*p)) ++p;
        if (!*p) break;
        char* e = nullptr;
        long a = strtol(p, &e, 10), b = a;
        p = e;
        if (*p == '-') { b = strtol(p + 1, &e, 10); p = e; }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (c >= 0 && CPU_ISSET((int)c, &allowed)) { CPU_SET((int)c, &want); ++n; }
    }
    if (n == 0) return PB200_OK;  // the cpuset of this process has no CPU near the device
    if (sched_setaffinity(0, sizeof want, &want) != 0) return PB200_OK;
    if (n_cpus_out) *n_cpus_out = n;
    return PB200_OK;
}

int pb200_host_alloc(uint64_t bytes, void** out) {
    if (!out) return set_error(PB200_ERR_INVALID, "null out");
    PB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return PB200_OK;
}
int pb200_host_free(void* p) {
    if (p) PB_CUDA(cudaFreeHost(p));
    return PB200_OK;
}
int pb200_device_alloc(pb200_ctx* ctx, uint64_t bytes, void** out) {
    PB_DEVICE(ctx);
    PB_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return PB200_OK;
}
int pb200_device_free(pb200_ctx* ctx, void* p) {
    PB_DEVICE(ctx);
    if (p) {
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
        PB_CUDA(cudaFree(p));
    }
    return PB200_OK;
}
int pb200_memcpy_h2d(pb200_ctx* ctx, void* d, const void* s, uint64_t bytes) {
    PB_DEVICE(ctx);
    PB_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
}
int pb200_memcpy_d2h(pb200_ctx* ctx, void* d, const void* s, uint64_t bytes) {
    PB_DEVICE(ctx);
    PB_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
}
int pb200_memcpy_d2d(pb200_ctx* ctx, void* d, const void* s, uint64_t bytes) {
    PB_DEVICE(ctx);
    PB_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return PB200_OK;
}
int pb200_memset_device(pb200_ctx* ctx, void* d, int value, uint64_t bytes) {
    PB_DEVICE(ctx);
    PB_CUDA(cudaMemsetAsync(d, value, bytes, ctx->stream));
    return PB200_OK;
}

// ---- PointLayout ---------------------------------------------------------------------------------

uint64_t pb200_dtype_size(uint32_t dtype, uint64_t extra_size) {  // point_layout.rs:72-97
    switch (dtype) {
        case PB200_U8: case PB200_I8: return 1;
        case PB200_U16: case PB200_I16: return 2;
        case PB200_U32: case PB200_I32: case PB200_F32: return 4;
        case PB200_U64: case PB200_I64: case PB200_F64: return 8;
        case PB200_VEC3U8: return 3;
        case PB200_VEC3U16: return 6;
        case PB200_VEC3I32: case PB200_VEC3F32: return 12;
        case PB200_VEC3F64: return 24;
        case PB200_VEC4U8: return 4;
        case PB200_BYTEARRAY: case PB200_CUSTOM: return extra_size;
        default: return 0;
    }
}

uint64_t pb200_dtype_min_alignment(uint32_t dtype, uint64_t extra_align) {  // point_layout.rs:100-126
    switch (dtype) {
        case PB200_U16: case PB200_I16: case PB200_VEC3U16: return 2;
        case PB200_U32: case PB200_I32: case PB200_F32: case PB200_VEC3I32: case PB200_VEC3F32: return 4;
        case PB200_U64: case PB200_I64: case PB200_F64: case PB200_VEC3F64: return 8;
        case PB200_CUSTOM: return extra_align ? extra_align : 1;
        default: return 1;
    }
}

int pb200_layout_create(pb200_layout** out) {
    if (!out) return set_error(PB200_ERR_INVALID, "null out");
    *out = new pb200_layout();
    return PB200_OK;
}

int pb200_layout_index_by_name(const pb200_layout* l, const char* name) {
    if (!l || !name) return -1;
    for (size_t i = 0; i < l->attrs.size(); ++i)
        if (strcmp(l->attrs[i].name, name) == 0) return (int)i;
    return -1;
}

int pb200_layout_index_of(const pb200_layout* l, const char* name, uint32_t dtype) {
    if (!l || !name) return -1;
    for (size_t i = 0; i < l->attrs.size(); ++i)
        if (strcmp(l->attrs[i].name, name) == 0 && l->attrs[i].dtype == dtype) return (int)i;
    return -1;
}

int pb200_layout_add_attribute(pb200_layout* l, const char* name, uint32_t dtype, uint64_t extra_size,
                               uint64_t extra_align, uint64_t packed_n) {
    if (!l || !name) return set_error(PB200_ERR_INVALID, "add_attribute: null argument");
    if (dtype > PB200_CUSTOM) return set_error(PB200_ERR_INVALID, "add_attribute: unknown dtype %u", dtype);
    if (strlen(name) >= PB200_MAX_NAME) return set_error(PB200_ERR_INVALID, "attribute name too long");
    if (l->attrs.size() >= PB200_MAX_ATTRIBUTES) return set_error(PB200_ERR_INVALID, "too many attributes");
    if (pb200_layout_index_by_name(l, name) >= 0)
        return set_error(PB200_ERR_DUPLICATE_ATTR, "Point attribute %s is already present in this PointLayout!", name);
    uint64_t min_align = pb200_dtype_min_alignment(dtype, extra_align);
    uint64_t field_align = packed_n ? (packed_n < min_align ? packed_n : min_align) : min_align;
    uint64_t next = l->attrs.empty() ? 0 : l->attrs.back().offset + l->attrs.back().size;
    uint64_t offset = align_to(next, field_align);
    uint64_t new_max = packed_n ? (packed_n < l->align ? packed_n : l->align) : (l->align > min_align ? l->align : min_align);
    pb200_attr a;
    memset(&a, 0, sizeof a);
    strncpy(a.name, name, PB200_MAX_NAME - 1);
    a.dtype = dtype;
    a.extra_size = extra_size;
    a.extra_align = extra_align;
    a.offset = offset;
    a.size = pb200_dtype_size(dtype, extra_size);
    l->attrs.push_back(a);
    uint64_t end = offset + a.size;
    uint64_t unaligned = l->size > end ? l->size : end;
    l->size = align_to(unaligned, new_max);
    l->align = new_max;
    return PB200_OK;
}

int pb200_layout_from_members_and_alignment(const pb200_attr* members, uint32_t n, uint64_t type_alignment,
                                            pb200_layout** out) {
    if (!out || (n && !members)) return set_error(PB200_ERR_INVALID, "null argument");
    if (n > PB200_MAX_ATTRIBUTES) return set_error(PB200_ERR_INVALID, "too many attributes");
    if (type_alignment == 0 || (type_alignment & (type_alignment - 1)))
        return set_error(PB200_ERR_INVALID, "Could not create memory layout for PointLayout (alignment %llu)",
                         (unsigned long long)type_alignment);
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = i + 1; j < n; ++j)
            if (strncmp(members[i].name, members[j].name, PB200_MAX_NAME) == 0)
                return set_error(PB200_ERR_DUPLICATE_ATTR, "All attributes must have unique names!");
    std::vector<std::pair<uint64_t, uint64_t>> ranges;
    for (uint32_t i = 0; i < n; ++i)
        ranges.push_back({members[i].offset, members[i].offset + pb200_dtype_size(members[i].dtype, members[i].extra_size)});
    for (size_t i = 1; i < ranges.size(); ++i) {  // stable insertion sort by start
        auto r = ranges[i];
        size_t j = i;
        while (j > 0 && ranges[j - 1].first > r.first) { ranges[j] = ranges[j - 1]; --j; }
        ranges[j] = r;
    }
    for (size_t i = 1; i < ranges.size(); ++i)
        if (ranges[i - 1].second > ranges[i].first)
            return set_error(PB200_ERR_OVERLAP, "All attributes must span non-overlapping memory regions!");
    pb200_layout* l = new pb200_layout();
    uint64_t unaligned = 0, max_off = 0;
    bool have = false;
    for (uint32_t i = 0; i < n; ++i) {
        pb200_attr a = members[i];
        a.name[PB200_MAX_NAME - 1] = 0;
        a.size = pb200_dtype_size(a.dtype, a.extra_size);
        if (!have || a.offset >= max_off) { max_off = a.offset; unaligned = a.offset + a.size; have = true; }
        l->attrs.push_back(a);
    }
    l->size = align_to(unaligned, type_alignment);
    l->align = type_alignment;
    *out = l;
    return PB200_OK;
}

int pb200_layout_clone(const pb200_layout* l, pb200_layout** out) {
    if (!l || !out) return set_error(PB200_ERR_INVALID, "null argument");
    *out = new pb200_layout(*l);
    return PB200_OK;
}
uint32_t pb200_layout_num_attributes(const pb200_layout* l) { return l ? (uint32_t)l->attrs.size() : 0; }
int pb200_layout_get_attribute(const pb200_layout* l, uint32_t i, pb200_attr* out) {
    if (!l || !out || i >= l->attrs.size()) return set_error(PB200_ERR_INVALID, "attribute index out of bounds");
    *out = l->attrs[i];
    return PB200_OK;
}
uint64_t pb200_layout_size_of_point_entry(const pb200_layout* l) { return l ? l->size : 0; }
uint64_t pb200_layout_alignment(const pb200_layout* l) { return l ? l->align : 0; }

int pb200_layout_equal(const pb200_layout* a, const pb200_layout* b) {
    if (!a || !b) return 0;
    if (a->attrs.size() != b->attrs.size() || a->size != b->size || a->align != b->align) return 0;
    for (size_t i = 0; i < a->attrs.size(); ++i) {
        const pb200_attr &x = a->attrs[i], &y = b->attrs[i];
        if (strcmp(x.name, y.name) != 0 || !dtype_equal(x, y) || x.offset != y.offset || x.size != y.size) return 0;
    }
    return 1;
}

int pb200_layout_compare_without_offsets(const pb200_layout* a, const pb200_layout* b) {
    if (!a || !b || a->attrs.size() != b->attrs.size()) return 0;
    for (const auto& x : a->attrs) {
        int j = pb200_layout_index_by_name(b, x.name);
        if (j < 0 || !dtype_equal(x, b->attrs[(size_t)j])) return 0;
    }
    return 1;
}

void pb200_layout_destroy(pb200_layout* l) { delete l; }

// ---- LAS layouts (pasture-io/src/las/las_layout.rs:64-125, las_types.rs) ------------------------------

struct LasFormat { bool extended, gps, color, nir, waveform; };
static int las_format_of(int n, LasFormat* f) {
    if (n < 0 || n > 10) return set_error(PB200_ERR_INVALID, "Unsupported LAS point format %d", n);
    f->extended = n >= 6;
    f->gps = (n == 1 || n == 3 || n == 4 || n == 5 || n >= 6);
    f->color = (n == 2 || n == 3 || n == 5 || n == 7 || n == 8 || n == 10);
    f->nir = (n == 8 || n == 10);
    f->waveform = (n == 4 || n == 5 || n == 9 || n == 10);
    return PB200_OK;
}
static void add_tail(pb200_layout* l, const LasFormat& f) {
    if (f.gps) pb200_layout_add_attribute(l, "GpsTime", PB200_F64, 0, 0, 1);
    if (f.color) pb200_layout_add_attribute(l, "ColorRGB", PB200_VEC3U16, 0, 0, 1);
    if (f.nir) pb200_layout_add_attribute(l, "NIR", PB200_U16, 0, 0, 1);
    if (f.waveform) {
        pb200_layout_add_attribute(l, "WavePacketDescriptorIndex", PB200_U8, 0, 0, 1);
        pb200_layout_add_attribute(l, "WaveformDataOffset", PB200_U64, 0, 0, 1);
        pb200_layout_add_attribute(l, "WaveformPacketSize", PB200_U32, 0, 0, 1);
        pb200_layout_add_attribute(l, "ReturnPointWaveformLocation", PB200_F32, 0, 0, 1);
        pb200_layout_add_attribute(l, "WaveformParameters", PB200_VEC3F32, 0, 0, 1);
    }
}

int pb200_las_raw_layout(int format, pb200_layout** out) {
    LasFormat f;
    PB_TRY(las_format_of(format, &f));
    pb200_layout* l = new pb200_layout();
    pb200_layout_add_attribute(l, "LASLocalPosition", PB200_VEC3I32, 0, 0, 1);
    pb200_layout_add_attribute(l, "Intensity", PB200_U16, 0, 0, 1);
    if (f.extended) pb200_layout_add_attribute(l, "LASExtendedFlags", PB200_U16, 0, 0, 1);
    else pb200_layout_add_attribute(l, "LASBasicFlags", PB200_U8, 0, 0, 1);
    pb200_layout_add_attribute(l, "Classification", PB200_U8, 0, 0, 1);
    if (f.extended) {
        pb200_layout_add_attribute(l, "UserData", PB200_U8, 0, 0, 1);
        pb200_layout_add_attribute(l, "ScanAngle", PB200_I16, 0, 0, 1);
    } else {
        pb200_layout_add_attribute(l, "ScanAngleRank", PB200_I8, 0, 0, 1);
        pb200_layout_add_attribute(l, "UserData", PB200_U8, 0, 0, 1);
    }
    pb200_layout_add_attribute(l, "PointSourceID", PB200_U16, 0, 0, 1);
    add_tail(l, f);
    *out = l;
    return PB200_OK;
}

int pb200_las_default_layout(int format, pb200_layout** out) {
    LasFormat f;
    PB_TRY(las_format_of(format, &f));
    pb200_layout* l = new pb200_layout();
    pb200_layout_add_attribute(l, "Position3D", PB200_VEC3F64, 0, 0, 1);
    pb200_layout_add_attribute(l, "Intensity", PB200_U16, 0, 0, 1);
    pb200_layout_add_attribute(l, "ReturnNumber", PB200_U8, 0, 0, 1);
    pb200_layout_add_attribute(l, "NumberOfReturns", PB200_U8, 0, 0, 1);
    if (f.extended) {
        pb200_layout_add_attribute(l, "ClassificationFlags", PB200_U8, 0, 0, 1);
        pb200_layout_add_attribute(l, "ScannerChannel", PB200_U8, 0, 0, 1);
    }
    pb200_layout_add_attribute(l, "ScanDirectionFlag", PB200_U8, 0, 0, 1);
    pb200_layout_add_attribute(l, "EdgeOfFlightLine", PB200_U8, 0, 0, 1);
    pb200_layout_add_attribute(l, "Classification", PB200_U8, 0, 0, 1);
    if (f.extended) {
        pb200_layout_add_attribute(l, "UserData", PB200_U8, 0, 0, 1);
        pb200_layout_add_attribute(l, "ScanAngle", PB200_I16, 0, 0, 1);
    } else {
        pb200_layout_add_attribute(l, "ScanAngleRank", PB200_I8, 0, 0, 1);
        pb200_layout_add_attribute(l, "UserData", PB200_U8, 0, 0, 1);
    }
    pb200_layout_add_attribute(l, "PointSourceID", PB200_U16, 0, 0, 1);
    add_tail(l, f);
    *out = l;
    return PB200_OK;
}

uint64_t pb200_expand_bits_by_3(uint64_t val) {  // math/bitmanip.rs:2-10
    val &= 0x1FFFFFull;
    val = (val | (val << 32)) & 0x00FF00000000FFFFull;
    val = (val | (val << 16)) & 0x00FF0000FF0000FFull;
    val = (val | (val << 8)) & 0xF00F00F00F00F00Full;
    val = (val | (val << 4)) & 0x30C30C30C30C30C3ull;
    val = (val | (val << 2)) & 0x1249249249249249ull;
    return val;
}

}  // extern "C"
