// voxel.cu -- voxel-grid downsampling (pasture-algorithms/src/voxel_grid.rs:109-165) as a sort-based GPU pipeline.
//
//   reference                                   here
//   calculate_bounds (:124)                     K5 reduction
//   create_markers_for_axis (:54-79)            host: the same running sum (sequential by definition), uploaded
//   find_leaf per point (:22-51, linear scan)   K7 voxel_key_kernel: O(1) guess + local search on the exact markers,
//                                               nearest-marker fix-up, packed key (ix,iy,iz)
//   sorted Vec<Voxel> with insert (:141-152)    K8 stable LSD radix sort of (key, point index) over the significant bits
//   per-voxel attribute reduction (:168-689)    K9 one thread per voxel walks its points IN INPUT ORDER (stable sort),
//                                               so f64 sums round exactly like the reference's sequential loops
// Output order = lexicographic (ix,iy,iz) = ascending packed key.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cfloat>
#include <climits>

#include "internal.h"

struct pb200_result_buffer {
    pb200_ctx* ctx = nullptr;
    pb200_layout layout;
    int32_t kind = 0, memspace = 0;
    uint64_t len = 0;
    void* aos = nullptr;
    std::vector<void*> columns;
    void* d_packed_keys = nullptr;  // device, one packed u64 per voxel (unpacked on demand)
    unsigned bits_y = 0, bits_z = 0;
};

namespace pb200 {

struct AxisGrid {
    const double* markers[3];
    unsigned long long n[3];
    double bmin[3], inv_leaf[3];
    unsigned bits_y, bits_z;
};

__device__ __forceinline__ unsigned long long leaf_index(double p, const double* __restrict__ m, unsigned long long n,
                                                         double bmin, double inv_leaf) {
    if (n == 0) return 0;  // voxel_grid.rs:31 `!markers.is_empty()`
    // first index with !(m[i] < p): guess from the regular spacing, then walk on the exact (running-sum) markers
    double g = (p - bmin) * inv_leaf;
    long long k = (g > 1.0) ? ((g < 9.0e18) ? (long long)g - 1 : (long long)n - 1) : 0;
    if (k > (long long)n - 1) k = (long long)n - 1;
    while (k < (long long)n - 1 && m[k] < p) ++k;
    while (k > 0 && !(m[k - 1] < p)) --k;
    unsigned long long i = (unsigned long long)k;
    // clamp to the better fitting marker: [i] or [i-1] (voxel_grid.rs:41-49)
    if (i > 0 && __dsub_rn(p, m[i - 1]) < __dsub_rn(m[i], p)) --i;
    return i;
}

__global__ void __launch_bounds__(256) voxel_key_kernel(const uint8_t* __restrict__ pos_base, unsigned long long stride,
                                                        unsigned long long n, AxisGrid g,
                                                        unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const double* p = reinterpret_cast<const double*>(pos_base + i * stride);
        const unsigned long long ix = leaf_index(p[0], g.markers[0], g.n[0], g.bmin[0], g.inv_leaf[0]);
        const unsigned long long iy = leaf_index(p[1], g.markers[1], g.n[1], g.bmin[1], g.inv_leaf[1]);
        const unsigned long long iz = leaf_index(p[2], g.markers[2], g.n[2], g.bmin[2], g.inv_leaf[2]);
        keys[i] = (((ix << g.bits_y) | iy) << g.bits_z) | iz;
        idx[i] = (uint32_t)i;
    }
}

__global__ void __launch_bounds__(256) head_flags_kernel(const unsigned long long* __restrict__ keys, unsigned long long n,
                                                         uint32_t* __restrict__ flags) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// seg[i] = exclusive scan of the head flags + own flag - 1 = voxel id of sorted position i
__global__ void __launch_bounds__(256) segment_starts_kernel(const uint32_t* __restrict__ flags,
                                                             const uint32_t* __restrict__ excl, unsigned long long n,
                                                             uint32_t* __restrict__ starts,
                                                             const unsigned long long* __restrict__ keys,
                                                             unsigned long long* __restrict__ voxel_keys) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        if (flags[i]) {
            starts[excl[i]] = (uint32_t)i;
            voxel_keys[excl[i]] = keys[i];
        }
}

template <class T>
__device__ __forceinline__ T ld_attr(const uint8_t* p, bool aligned) {
    if (aligned) return *reinterpret_cast<const T*>(p);
    T v;
    uint8_t* b = reinterpret_cast<uint8_t*>(&v);
#pragma unroll
    for (int k = 0; k < (int)sizeof(T); ++k) b[k] = p[k];
    return v;
}

__device__ __forceinline__ unsigned long long f64_as_u64_sat(double v) {  // Rust `as u64`
    if (!(v > 0.0)) return 0;
    if (v >= 18446744073709551616.0) return ULLONG_MAX;
    return __double2ull_rz(v);
}

enum ReduceKind { R_MEAN_VEC_F64 = 0, R_MEAN_U16, R_MEAN_VEC_U16, R_MEAN_VEC_F32, R_MODE, R_MODE_BOOL, R_MAX_U8, R_MAX_F64, R_MAX_U64 };

struct ReduceArgs {
    const uint32_t* starts;     // V+1 entries (starts[V] = N)
    const uint32_t* sorted_idx; // N
    unsigned long long n_voxels;
    const uint8_t* src;         // attribute of point 0
    unsigned long long src_stride;
    uint8_t* dst;               // column of the result (SoA staging), element size = dst_size
    uint32_t dst_size;
    int src_aligned;
};

// S = source scalar (component) type. One thread per voxel; points are visited in input order.
template <class S, int KIND>
__global__ void __launch_bounds__(128) voxel_reduce_kernel(ReduceArgs a) {
    const unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.n_voxels) return;
    const uint32_t b = a.starts[v], e = a.starts[v + 1];
    const bool al = a.src_aligned != 0;
    uint8_t* out = a.dst + v * a.dst_size;
    if constexpr (KIND == R_MEAN_VEC_F64 || KIND == R_MEAN_VEC_U16 || KIND == R_MEAN_VEC_F32) {
        double sx = 0.0, sy = 0.0, sz = 0.0;  // voxel_grid.rs:339-379
        for (uint32_t k = b; k < e; ++k) {
            const uint8_t* p = a.src + (unsigned long long)a.sorted_idx[k] * a.src_stride;
            sx = __dadd_rn(sx, (double)ld_attr<S>(p, al));
            sy = __dadd_rn(sy, (double)ld_attr<S>(p + sizeof(S), al));
            sz = __dadd_rn(sz, (double)ld_attr<S>(p + 2 * sizeof(S), al));
        }
        const double cnt = (double)(e - b);  // :382-386
        const double cx = sx / cnt, cy = sy / cnt, cz = sz / cnt;
        if constexpr (KIND == R_MEAN_VEC_F64) {
            double r[3] = {cx, cy, cz};
            memcpy(out, r, 24);
        } else if constexpr (KIND == R_MEAN_VEC_U16) {  // `as u16` :597
            uint16_t r[3];
            const double c[3] = {cx, cy, cz};
            for (int j = 0; j < 3; ++j) { unsigned long long u = f64_as_u64_sat(c[j]); r[j] = (uint16_t)(u > 65535ull ? 65535ull : u); }
            memcpy(out, r, 6);
        } else {  // `as f32` :674
            float r[3] = {(float)cx, (float)cy, (float)cz};
            memcpy(out, r, 12);
        }
    } else if constexpr (KIND == R_MEAN_U16) {  // :391-439, `as u16` :474,:612
        double s = 0.0;
        for (uint32_t k = b; k < e; ++k)
            s = __dadd_rn(s, (double)ld_attr<S>(a.src + (unsigned long long)a.sorted_idx[k] * a.src_stride, al));
        unsigned long long u = f64_as_u64_sat(s / (double)(e - b));
        uint16_t r = (uint16_t)(u > 65535ull ? 65535ull : u);
        memcpy(out, &r, 2);
    } else if constexpr (KIND == R_MODE || KIND == R_MODE_BOOL) {
        // most common value (:218-329); ties: the reference picks in HashMap order (nondeterministic), here the smallest
        long long best = 0;
        uint32_t best_cnt = 0;
        for (uint32_t k = b; k < e; ++k) {
            const long long val = (long long)ld_attr<S>(a.src + (unsigned long long)a.sorted_idx[k] * a.src_stride, al);
            bool seen = false;  // count each distinct value once, at its first occurrence
            for (uint32_t j = b; j < k && !seen; ++j)
                seen = (long long)ld_attr<S>(a.src + (unsigned long long)a.sorted_idx[j] * a.src_stride, al) == val;
            if (seen) continue;
            uint32_t cnt = 1;
            for (uint32_t j = k + 1; j < e; ++j)
                cnt += (long long)ld_attr<S>(a.src + (unsigned long long)a.sorted_idx[j] * a.src_stride, al) == val ? 1u : 0u;
            if (cnt > best_cnt || (cnt == best_cnt && val < best)) { best = val; best_cnt = cnt; }
        }
        if constexpr (KIND == R_MODE_BOOL) { uint8_t r = best != 0 ? 1 : 0; memcpy(out, &r, 1); }  // :527,:540
        else memcpy(out, &best, a.dst_size);  // `as u8/i8/i16/u16`: low bytes
    } else {  // max-pool starting from 0.0 (:168-215)
        double cur = 0.0;
        for (uint32_t k = b; k < e; ++k) {
            const double x = (double)ld_attr<S>(a.src + (unsigned long long)a.sorted_idx[k] * a.src_stride, al);
            if (x > cur) cur = x;
        }
        if constexpr (KIND == R_MAX_U8) { unsigned long long u = f64_as_u64_sat(cur); uint8_t r = (uint8_t)(u > 255ull ? 255ull : u); memcpy(out, &r, 1); }
        else if constexpr (KIND == R_MAX_F64) memcpy(out, &cur, 8);
        else { unsigned long long r = f64_as_u64_sat(cur); memcpy(out, &r, 8); }
    }
}

// ---- sort-based mode for crowded voxels -------------------------------------------------------------------------
// The per-thread mode above is O(points-in-voxel^2): fine for the usual handful of points per voxel, hopeless for a
// voxel that swallows a large part of the cloud.  When the most crowded voxel holds more than MODE_THREAD_LIMIT points
// the mode attributes switch to: composite key (voxel id << 16 | biased value) -> radix sort -> run lengths ->
// one 64-bit atomicMax per run on (length << 16 | 65535 - biased value): the longest run wins, ties go to the smallest
// value (the same rule as the per-thread path).
constexpr uint32_t MODE_THREAD_LIMIT = 64;

__global__ void __launch_bounds__(256) max_occupancy_kernel(const uint32_t* __restrict__ starts, unsigned long long n_voxels,
                                                            uint32_t* __restrict__ out) {
    uint32_t m = 0;
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_voxels; v += step) {
        const uint32_t c = starts[v + 1] - starts[v];
        m = c > m ? c : m;
    }
    for (int o = 16; o > 0; o >>= 1) { const uint32_t x = __shfl_xor_sync(0xffffffffu, m, o); m = x > m ? x : m; }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

template <class S>
__global__ void __launch_bounds__(256) mode_keys_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ excl,
                                                        const uint32_t* __restrict__ sorted_idx, unsigned long long n,
                                                        const uint8_t* __restrict__ src, unsigned long long stride, int aligned,
                                                        unsigned long long* __restrict__ keys) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const unsigned long long voxel = (unsigned long long)excl[i] + flags[i] - 1ull;
        const long long val = (long long)ld_attr<S>(src + (unsigned long long)sorted_idx[i] * stride, aligned != 0);
        constexpr long long bias = (S(-1) < S(0)) ? 32768 : 0;  // keeps signed values ordered as unsigned 16-bit fields
        keys[i] = (voxel << 16) | (unsigned long long)((val + bias) & 0xFFFF);
    }
}

__global__ void __launch_bounds__(256) run_heads_kernel(const unsigned long long* __restrict__ keys, unsigned long long n,
                                                        uint32_t* __restrict__ head_idx) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        head_idx[i] = (i == 0 || keys[i] != keys[i - 1]) ? (uint32_t)i : 0u;
}

__global__ void __launch_bounds__(256) run_vote_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ head_of,
                                                       unsigned long long n, unsigned long long* __restrict__ best) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        if (i + 1 < n && keys[i + 1] == keys[i]) continue;  // only the last element of a run votes
        const unsigned long long len = i - head_of[i] + 1ull;
        const unsigned long long voxel = keys[i] >> 16, biased = keys[i] & 0xFFFFull;
        atomicMax(&best[voxel], (len << 16) | (65535ull - biased));
    }
}

__global__ void __launch_bounds__(256) mode_decode_kernel(const unsigned long long* __restrict__ best, unsigned long long n_voxels,
                                                          uint8_t* __restrict__ dst, uint32_t dst_size, int as_bool, long long bias) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_voxels; v += step) {
        const long long val = (long long)(65535ull - (best[v] & 0xFFFFull)) - bias;
        if (as_bool) dst[v] = val != 0 ? 1 : 0;
        else memcpy(dst + v * dst_size, &val, dst_size);
    }
}

struct MaxU32 {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

struct Rule { const char* name; uint32_t dtype; ReduceKind kind; };
static const Rule RULES[] = {  // voxel_grid.rs:461-679, source order
    {"Position3D", PB200_VEC3F64, R_MEAN_VEC_F64}, {"Intensity", PB200_U16, R_MEAN_U16},
    {"ReturnNumber", PB200_U8, R_MODE}, {"NumberOfReturns", PB200_U8, R_MODE},
    {"ClassificationFlags", PB200_U8, R_MAX_U8}, {"ScannerChannel", PB200_U8, R_MODE},
    {"ScanDirectionFlag", PB200_U8, R_MODE_BOOL}, {"EdgeOfFlightLine", PB200_U8, R_MODE_BOOL},
    {"Classification", PB200_U8, R_MODE}, {"ScanAngleRank", PB200_I8, R_MODE}, {"ScanAngle", PB200_I16, R_MODE},
    {"UserData", PB200_U8, R_MODE}, {"PointSourceID", PB200_U16, R_MODE}, {"ColorRGB", PB200_VEC3U16, R_MEAN_VEC_U16},
    {"GpsTime", PB200_F64, R_MAX_F64}, {"NIR", PB200_U16, R_MEAN_U16}, {"PointID", PB200_U64, R_MAX_U64},
    {"Normal", PB200_VEC3F32, R_MEAN_VEC_F32},
};
static const Rule* find_rule(const pb200_attr& a) {
    for (const Rule& r : RULES)
        if (strcmp(r.name, a.name) == 0 && r.dtype == a.dtype) return &r;
    return nullptr;
}

static void launch_reduce(const Rule& r, const ReduceArgs& a, cudaStream_t st) {
    const unsigned blocks = (unsigned)((a.n_voxels + 127) / 128);
    switch (r.kind) {
        case R_MEAN_VEC_F64: voxel_reduce_kernel<double, R_MEAN_VEC_F64><<<blocks, 128, 0, st>>>(a); break;
        case R_MEAN_U16: voxel_reduce_kernel<uint16_t, R_MEAN_U16><<<blocks, 128, 0, st>>>(a); break;
        case R_MEAN_VEC_U16: voxel_reduce_kernel<uint16_t, R_MEAN_VEC_U16><<<blocks, 128, 0, st>>>(a); break;
        case R_MEAN_VEC_F32: voxel_reduce_kernel<float, R_MEAN_VEC_F32><<<blocks, 128, 0, st>>>(a); break;
        case R_MODE:
            if (r.dtype == PB200_U8) voxel_reduce_kernel<uint8_t, R_MODE><<<blocks, 128, 0, st>>>(a);
            else if (r.dtype == PB200_I8) voxel_reduce_kernel<int8_t, R_MODE><<<blocks, 128, 0, st>>>(a);
            else if (r.dtype == PB200_I16) voxel_reduce_kernel<int16_t, R_MODE><<<blocks, 128, 0, st>>>(a);
            else voxel_reduce_kernel<uint16_t, R_MODE><<<blocks, 128, 0, st>>>(a);
            break;
        case R_MODE_BOOL: voxel_reduce_kernel<uint8_t, R_MODE_BOOL><<<blocks, 128, 0, st>>>(a); break;
        case R_MAX_U8: voxel_reduce_kernel<uint8_t, R_MAX_U8><<<blocks, 128, 0, st>>>(a); break;
        case R_MAX_F64: voxel_reduce_kernel<double, R_MAX_F64><<<blocks, 128, 0, st>>>(a); break;
        case R_MAX_U64: voxel_reduce_kernel<unsigned long long, R_MAX_U64><<<blocks, 128, 0, st>>>(a); break;
    }
    g_launches++;
}

static unsigned bits_for(unsigned long long count) {  // bits needed for indices 0..count-1 (at least 1)
    unsigned b = 1;
    while ((1ull << b) < count) ++b;
    return b;
}

struct DeviceBuf {  // RAII for temporaries
    void* p = nullptr;
    ~DeviceBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_voxelgrid_filter(pb200_ctx* ctx, const pb200_buffer_desc* src, double lx, double ly, double lz,
                           const pb200_layout* dst_layout, int32_t dst_kind, int32_t dst_memspace,
                           pb200_result_buffer** out) {
    if (!ctx || !dst_layout || !out) return set_error(PB200_ERR_INVALID, "null argument");
    *out = nullptr;
    PB_TRY(validate_desc(src, "source buffer"));
    PB_TRY(ensure_device(ctx));
    if (dst_kind != PB200_INTERLEAVED && dst_kind != PB200_COLUMNAR) return set_error(PB200_ERR_INVALID, "bad dst_kind");
    const int pi = pb200_layout_index_of(src->layout, "Position3D", PB200_VEC3F64);
    if (pi < 0)  // voxel_grid.rs:116-121
        return set_error(PB200_ERR_ATTR_NOT_FOUND, "The PointBuffer does not have the attribute attributes::POSITION_3D which is needed for the creation of the voxel grid.");
    const double leaf[3] = {lx, ly, lz};
    for (int c = 0; c < 3; ++c)
        if (!(leaf[c] > 0.0)) return set_error(PB200_ERR_INVALID, "leaf sizes must be positive (the reference would not terminate)");
    if (src->len == 0) return set_error(PB200_ERR_INVALID, "calculate_bounds returned None (empty buffer): Option::unwrap() panics (voxel_grid.rs:124)");
    if (src->len > 0xFFFFFFFFull) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^32-1 points per call");
    // target layout rules (:452-459, :461-687)
    static const char* WAVE[] = {"WaveformDataOffset", "WaveformPacketSize", "WaveformParameters", "WavePacketDescriptorIndex", "ReturnPointWaveformLocation"};
    static const uint32_t WAVE_T[] = {PB200_U64, PB200_U32, PB200_VEC3F32, PB200_U8, PB200_F32};
    for (int w = 0; w < 5; ++w)
        if (pb200_layout_index_of(dst_layout, WAVE[w], WAVE_T[w]) >= 0) return set_error(PB200_ERR_UNSUPPORTED, "Waveform data currently not supported!");
    std::vector<const Rule*> rules;
    std::vector<int> src_idx;
    for (const auto& a : dst_layout->attrs) {
        const Rule* r = find_rule(a);
        if (!r) return set_error(PB200_ERR_UNSUPPORTED, "attribute is non-standard which is not supported currently: %s", a.name);
        int si = pb200_layout_index_of(src->layout, r->name, r->dtype);
        if (si < 0) return set_error(PB200_ERR_ATTR_NOT_FOUND, "source buffer has no attribute %s with the default datatype", r->name);
        rules.push_back(r);
        src_idx.push_back(si);
    }
    const uint64_t n = src->len;
    cudaStream_t st = ctx->stream;

    // ---- source attribute streams on the device --------------------------------------------------------------
    std::vector<DeviceBuf> staged(src->layout->attrs.size() + 1);
    const uint8_t* d_aos = nullptr;
    std::vector<const uint8_t*> d_cols(src->layout->attrs.size(), nullptr);
    auto need = [&](int idx) -> int {
        if (src->memspace == PB200_DEVICE) {
            if (src->kind == PB200_INTERLEAVED) d_aos = (const uint8_t*)src->aos;
            else d_cols[(size_t)idx] = (const uint8_t*)src->columns[idx];
            return PB200_OK;
        }
        if (src->kind == PB200_INTERLEAVED) {
            if (!d_aos) {
                PB_CUDA(staged.back().alloc((size_t)(n * src->layout->size)));
                PB_CUDA(cudaMemcpyAsync(staged.back().p, src->aos, (size_t)(n * src->layout->size), cudaMemcpyHostToDevice, st));
                d_aos = (const uint8_t*)staged.back().p;
            }
        } else if (!d_cols[(size_t)idx]) {
            const size_t bytes = (size_t)(n * src->layout->attrs[(size_t)idx].size);
            PB_CUDA(staged[(size_t)idx].alloc(bytes));
            PB_CUDA(cudaMemcpyAsync(staged[(size_t)idx].p, src->columns[idx], bytes, cudaMemcpyHostToDevice, st));
            d_cols[(size_t)idx] = (const uint8_t*)staged[(size_t)idx].p;
        }
        return PB200_OK;
    };
    PB_TRY(need(pi));
    for (int si : src_idx) PB_TRY(need(si));
    auto attr_ptr = [&](int idx, uint64_t* stride) -> const uint8_t* {
        const pb200_attr& a = src->layout->attrs[(size_t)idx];
        if (src->kind == PB200_INTERLEAVED) { *stride = src->layout->size; return d_aos + a.offset; }
        *stride = a.size;
        return d_cols[(size_t)idx];
    };

    // ---- bounds + markers -------------------------------------------------------------------------------------
    pb200_buffer_desc dsrc = *src;  // device view for the bounds reduction
    std::vector<void*> dcol_ptrs(d_cols.size());
    for (size_t i = 0; i < d_cols.size(); ++i) dcol_ptrs[i] = (void*)d_cols[i];
    dsrc.memspace = PB200_DEVICE;
    dsrc.aos = (void*)d_aos;
    dsrc.columns = dcol_ptrs.data();
    double bmin[3], bmax[3];
    int some = 0;
    PB_TRY(pb200_calculate_bounds(ctx, &dsrc, bmin, bmax, &some));
    if (!some) return set_error(PB200_ERR_INVALID, "calculate_bounds returned None");
    std::vector<double> markers[3];
    for (int c = 0; c < 3; ++c) {  // voxel_grid.rs:63-77 running sum
        double cur = bmin[c];
        while (cur < bmax[c]) {
            cur += leaf[c];
            markers[c].push_back(cur);
            if (markers[c].size() > (1u << 21)) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^21 voxels along one axis");
        }
    }
    const unsigned bits_x = bits_for(markers[0].size()), bits_y = bits_for(markers[1].size()), bits_z = bits_for(markers[2].size());
    DeviceBuf d_markers;
    const size_t nm_total = markers[0].size() + markers[1].size() + markers[2].size();
    PB_CUDA(d_markers.alloc(nm_total * sizeof(double) + 8));
    AxisGrid grid;
    {
        size_t off = 0;
        for (int c = 0; c < 3; ++c) {
            grid.markers[c] = (const double*)d_markers.p + off;
            grid.n[c] = markers[c].size();
            grid.bmin[c] = bmin[c];
            grid.inv_leaf[c] = 1.0 / leaf[c];
            if (!markers[c].empty())
                PB_CUDA(cudaMemcpyAsync((double*)d_markers.p + off, markers[c].data(), markers[c].size() * sizeof(double), cudaMemcpyHostToDevice, st));
            off += markers[c].size();
        }
        grid.bits_y = bits_y;
        grid.bits_z = bits_z;
    }

    // ---- keys, sort, segments ---------------------------------------------------------------------------------
    DeviceBuf d_keys, d_keys2, d_idx, d_idx2, d_flags, d_excl, d_tmp;
    PB_CUDA(d_keys.alloc(n * 8)); PB_CUDA(d_keys2.alloc(n * 8));
    PB_CUDA(d_idx.alloc(n * 4)); PB_CUDA(d_idx2.alloc(n * 4));
    PB_CUDA(d_flags.alloc(n * 4)); PB_CUDA(d_excl.alloc(n * 4));
    uint64_t pstride = 0;
    const uint8_t* ppos = attr_ptr(pi, &pstride);
    if (((uintptr_t)ppos & 7) || (pstride & 7)) return set_error(PB200_ERR_UNSUPPORTED, "POSITION_3D must be 8-byte aligned in memory");
    const unsigned long long cap = (unsigned long long)ctx->sm_count * 16;
    const unsigned blocks = (unsigned)(((n + 255) / 256) < cap ? ((n + 255) / 256) : cap);
    voxel_key_kernel<<<blocks, 256, 0, st>>>(ppos, pstride, n, grid, (unsigned long long*)d_keys.p, (uint32_t*)d_idx.p);
    g_launches++;
    size_t tmp_bytes = 0, tmp2 = 0;
    const int end_bit = (int)(bits_x + bits_y + bits_z);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned long long*)d_keys.p, (unsigned long long*)d_keys2.p,
                                    (const uint32_t*)d_idx.p, (uint32_t*)d_idx2.p, (int)n, 0, end_bit, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, (const uint32_t*)d_flags.p, (uint32_t*)d_excl.p, (int)n, st);
    PB_CUDA(d_tmp.alloc(tmp_bytes > tmp2 ? tmp_bytes : tmp2));
    PB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, (const unsigned long long*)d_keys.p, (unsigned long long*)d_keys2.p,
                                            (const uint32_t*)d_idx.p, (uint32_t*)d_idx2.p, (int)n, 0, end_bit, st));
    g_launches += (uint64_t)((end_bit + 7) / 8) + 1;
    head_flags_kernel<<<blocks, 256, 0, st>>>((const unsigned long long*)d_keys2.p, n, (uint32_t*)d_flags.p);
    g_launches++;
    PB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp2, (const uint32_t*)d_flags.p, (uint32_t*)d_excl.p, (int)n, st));
    g_launches++;
    uint32_t last[2];
    PB_CUDA(cudaMemcpyAsync(&last[0], (uint32_t*)d_excl.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaMemcpyAsync(&last[1], (uint32_t*)d_flags.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    const uint64_t V = (uint64_t)last[0] + last[1];
    DeviceBuf d_starts, d_vkeys;
    PB_CUDA(d_starts.alloc((V + 1) * 4));
    PB_CUDA(d_vkeys.alloc(V * 8 + 8));
    segment_starts_kernel<<<blocks, 256, 0, st>>>((const uint32_t*)d_flags.p, (const uint32_t*)d_excl.p, n, (uint32_t*)d_starts.p,
                                                  (const unsigned long long*)d_keys2.p, (unsigned long long*)d_vkeys.p);
    g_launches++;
    const uint32_t n32 = (uint32_t)n;
    PB_CUDA(cudaMemcpyAsync((uint32_t*)d_starts.p + V, &n32, 4, cudaMemcpyHostToDevice, st));

    // the most crowded voxel decides how the mode attributes are reduced
    bool any_mode = false;
    for (const Rule* r : rules) any_mode = any_mode || r->kind == R_MODE || r->kind == R_MODE_BOOL;
    uint32_t max_occ = 0;
    DeviceBuf d_mode_keys, d_mode_keys2, d_head, d_head2, d_best, d_mode_tmp;
    size_t mode_tmp_bytes = 0;
    if (any_mode) {
        PB_CUDA(cudaMemsetAsync(d_tmp.p, 0, 4, st));  // d_tmp[0..3] doubles as the max counter (the sort is done)
        max_occupancy_kernel<<<blocks, 256, 0, st>>>((const uint32_t*)d_starts.p, V, (uint32_t*)d_tmp.p);
        g_launches++;
        PB_CUDA(cudaMemcpyAsync(&max_occ, d_tmp.p, 4, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        if (max_occ > MODE_THREAD_LIMIT) {
            PB_CUDA(d_mode_keys.alloc(n * 8)); PB_CUDA(d_mode_keys2.alloc(n * 8));
            PB_CUDA(d_head.alloc(n * 4)); PB_CUDA(d_head2.alloc(n * 4));
            PB_CUDA(d_best.alloc((V + 1) * 8));
            size_t t1 = 0, t2 = 0;
            cub::DeviceRadixSort::SortKeys(nullptr, t1, (const unsigned long long*)d_mode_keys.p, (unsigned long long*)d_mode_keys2.p, (int)n, 0, 64, st);
            cub::DeviceScan::InclusiveScan(nullptr, t2, (const uint32_t*)d_head.p, (uint32_t*)d_head2.p, MaxU32(), (int)n, st);
            mode_tmp_bytes = t1 > t2 ? t1 : t2;
            PB_CUDA(d_mode_tmp.alloc(mode_tmp_bytes));
        }
    }

    // ---- per-attribute reductions into a columnar staging result -----------------------------------------------
    pb200_result_buffer* res = new pb200_result_buffer();
    res->ctx = ctx;
    res->layout = *dst_layout;
    res->kind = dst_kind;
    res->memspace = dst_memspace;
    res->len = V;
    std::vector<void*> d_out(dst_layout->attrs.size(), nullptr);
    auto fail = [&](int rc) {
        for (void* p : d_out) if (p) cudaFree(p);
        if (res->d_packed_keys) cudaFree(res->d_packed_keys);
        delete res;
        return rc;
    };
    for (size_t a = 0; a < dst_layout->attrs.size(); ++a) {
        cudaError_t e = cudaMalloc(&d_out[a], (size_t)(V * dst_layout->attrs[a].size) + 16);
        if (e != cudaSuccess) return fail(cuda_error(e, "cudaMalloc(result column)"));
        ReduceArgs ra;
        uint64_t sstride = 0;
        ra.src = attr_ptr(src_idx[a], &sstride);
        ra.src_stride = sstride;
        ra.starts = (const uint32_t*)d_starts.p;
        ra.sorted_idx = (const uint32_t*)d_idx2.p;
        ra.n_voxels = V;
        ra.dst = (uint8_t*)d_out[a];
        ra.dst_size = (uint32_t)dst_layout->attrs[a].size;
        const uint64_t comp = pb200_dtype_size(is_cast_vec3(rules[a]->dtype) ? vec3_component(rules[a]->dtype) : rules[a]->dtype, 0);
        ra.src_aligned = (((uintptr_t)ra.src % comp) == 0 && (sstride % comp) == 0) ? 1 : 0;
        const bool is_mode = rules[a]->kind == R_MODE || rules[a]->kind == R_MODE_BOOL;
        if (is_mode && max_occ > MODE_THREAD_LIMIT) {
            const uint32_t dt = rules[a]->dtype;
            unsigned long long* mk = (unsigned long long*)d_mode_keys.p;
            const uint32_t *fl = (const uint32_t*)d_flags.p, *ex = (const uint32_t*)d_excl.p, *si2 = (const uint32_t*)d_idx2.p;
            if (dt == PB200_U8) mode_keys_kernel<uint8_t><<<blocks, 256, 0, st>>>(fl, ex, si2, n, ra.src, sstride, ra.src_aligned, mk);
            else if (dt == PB200_I8) mode_keys_kernel<int8_t><<<blocks, 256, 0, st>>>(fl, ex, si2, n, ra.src, sstride, ra.src_aligned, mk);
            else if (dt == PB200_I16) mode_keys_kernel<int16_t><<<blocks, 256, 0, st>>>(fl, ex, si2, n, ra.src, sstride, ra.src_aligned, mk);
            else mode_keys_kernel<uint16_t><<<blocks, 256, 0, st>>>(fl, ex, si2, n, ra.src, sstride, ra.src_aligned, mk);
            size_t tb = mode_tmp_bytes;
            cudaError_t e2 = cub::DeviceRadixSort::SortKeys(d_mode_tmp.p, tb, (const unsigned long long*)mk, (unsigned long long*)d_mode_keys2.p,
                                                            (int)n, 0, 16 + (int)bits_for(V + 1), st);
            if (e2 != cudaSuccess) return fail(cuda_error(e2, "mode sort"));
            run_heads_kernel<<<blocks, 256, 0, st>>>((const unsigned long long*)d_mode_keys2.p, n, (uint32_t*)d_head.p);
            tb = mode_tmp_bytes;
            e2 = cub::DeviceScan::InclusiveScan(d_mode_tmp.p, tb, (const uint32_t*)d_head.p, (uint32_t*)d_head2.p, MaxU32(), (int)n, st);
            if (e2 != cudaSuccess) return fail(cuda_error(e2, "mode scan"));
            cudaMemsetAsync(d_best.p, 0, (V + 1) * 8, st);
            run_vote_kernel<<<blocks, 256, 0, st>>>((const unsigned long long*)d_mode_keys2.p, (const uint32_t*)d_head2.p, n, (unsigned long long*)d_best.p);
            mode_decode_kernel<<<blocks, 256, 0, st>>>((const unsigned long long*)d_best.p, V, ra.dst, ra.dst_size, rules[a]->kind == R_MODE_BOOL ? 1 : 0,
                                                           (dt == PB200_I8 || dt == PB200_I16) ? 32768ll : 0ll);
            g_launches += 8;
        } else {
            launch_reduce(*rules[a], ra, st);
        }
    }
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(cuda_error(e, "voxel reduce"));
    }
    // packed voxel keys stay on the device; pb200_result_buffer_voxel_keys unpacks them on demand
    res->d_packed_keys = d_vkeys.p;
    d_vkeys.p = nullptr;
    res->bits_y = bits_y;
    res->bits_z = bits_z;
    cudaStreamSynchronize(st);  // temporaries (sorted keys, indices, staged inputs) are released on return
    // ---- hand the result over in the requested memory layout / space -------------------------------------------
    pb200_buffer_desc stage;
    stage.layout = dst_layout;
    stage.kind = PB200_COLUMNAR;
    stage.memspace = PB200_DEVICE;
    stage.len = V;
    stage.aos = nullptr;
    stage.columns = d_out.data();
    if (dst_kind == PB200_COLUMNAR && dst_memspace == PB200_DEVICE) {
        res->columns = d_out;
        *out = res;
        return PB200_OK;
    }
    // allocate the final storage
    if (dst_kind == PB200_INTERLEAVED) {
        const size_t bytes = (size_t)(V * dst_layout->size) + 16;
        if (dst_memspace == PB200_DEVICE) { if (cudaMalloc(&res->aos, bytes) != cudaSuccess) return fail(set_error(PB200_ERR_OOM, "out of device memory")); cudaMemsetAsync(res->aos, 0, bytes, st); }
        else { res->aos = calloc(1, bytes); if (!res->aos) return fail(set_error(PB200_ERR_OOM, "out of host memory")); }
    } else {
        res->columns.assign(dst_layout->attrs.size(), nullptr);
        for (size_t a = 0; a < dst_layout->attrs.size(); ++a) {
            res->columns[a] = calloc(1, (size_t)(V * dst_layout->attrs[a].size) + 16);
            if (!res->columns[a]) return fail(set_error(PB200_ERR_OOM, "out of host memory"));
        }
    }
    pb200_buffer_desc fin;
    fin.layout = dst_layout;
    fin.kind = dst_kind;
    fin.memspace = dst_memspace;
    fin.len = V;
    fin.aos = res->aos;
    fin.columns = res->columns.empty() ? nullptr : res->columns.data();
    int rc = PB200_OK;
    if (V) {
        pb200_converter* cv = nullptr;
        rc = pb200_converter_create(ctx, dst_layout, dst_layout, 1, &cv);  // identity mappings: columnar staging -> final
        if (rc == PB200_OK) rc = pb200_converter_convert_into_range(cv, &stage, 0, V, &fin, 0, V, nullptr);
        pb200_converter_destroy(cv);
        cudaStreamSynchronize(st);
    }
    for (void* p : d_out) if (p) cudaFree(p);
    if (rc < 0) {
        std::fill(d_out.begin(), d_out.end(), nullptr);
        pb200_result_buffer_destroy(res);
        return rc;
    }
    *out = res;
    return PB200_OK;
}

int pb200_result_buffer_desc(const pb200_result_buffer* r, pb200_buffer_desc* out) {
    if (!r || !out) return set_error(PB200_ERR_INVALID, "null argument");
    out->layout = &r->layout;
    out->kind = r->kind;
    out->memspace = r->memspace;
    out->len = r->len;
    out->aos = r->aos;
    out->columns = r->columns.empty() ? nullptr : const_cast<void**>(r->columns.data());
    return PB200_OK;
}

int pb200_result_buffer_voxel_keys(const pb200_result_buffer* r, uint64_t* keys_out) {
    if (!r || !keys_out) return set_error(PB200_ERR_INVALID, "null argument");
    if (r->len == 0) return PB200_OK;
    PB_TRY(ensure_device(r->ctx));
    std::vector<unsigned long long> packed(r->len);
    PB_CUDA(cudaMemcpyAsync(packed.data(), r->d_packed_keys, r->len * 8, cudaMemcpyDeviceToHost, r->ctx->stream));
    PB_CUDA(cudaStreamSynchronize(r->ctx->stream));
    for (uint64_t v = 0; v < r->len; ++v) {
        const unsigned long long k = packed[v];
        keys_out[3 * v + 2] = k & ((1ull << r->bits_z) - 1);
        keys_out[3 * v + 1] = (k >> r->bits_z) & ((1ull << r->bits_y) - 1);
        keys_out[3 * v + 0] = k >> (r->bits_z + r->bits_y);
    }
    return PB200_OK;
}

void pb200_result_buffer_destroy(pb200_result_buffer* r) {
    if (!r) return;
    if (r->ctx) cudaSetDevice(r->ctx->device);
    if (r->d_packed_keys) cudaFree(r->d_packed_keys);
    if (r->memspace == PB200_DEVICE) {
        if (r->aos) cudaFree(r->aos);
        for (void* p : r->columns) if (p) cudaFree(p);
    } else {
        free(r->aos);
        for (void* p : r->columns) free(p);
    }
    delete r;
}

}  // extern "C"
