// reduce.cu -- sweeps that reduce every point: AABB (pasture-algorithms/src/bounds.rs:11-85),
// per-attribute min/max (pasture-algorithms/src/minmax.rs:13-51) and 63-bit Morton codes
// (pasture-core/src/math/bitmanip.rs:2-10 as the bit spreader).
//
// All reductions run on order-preserving 64-bit keys (signed ints: flip the sign bit; floats: the
// usual sign-magnitude fix-up after widening to f64, NaN never enters) so that one warp-shuffle +
// shared-memory + global-atomic ladder serves every dtype.
#include <cfloat>

#include "internal.h"

namespace pb200 {

__device__ __forceinline__ unsigned long long key_of_f64(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
static inline double f64_of_key(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

template <class T> struct KeyOf;
#define PB_KEY_UNSIGNED(T) template <> struct KeyOf<T> { static constexpr bool fp = false; \
    __device__ static unsigned long long key(T v) { return (unsigned long long)v; } };
#define PB_KEY_SIGNED(T) template <> struct KeyOf<T> { static constexpr bool fp = false; \
    __device__ static unsigned long long key(T v) { return (unsigned long long)(long long)v ^ 0x8000000000000000ull; } };
PB_KEY_UNSIGNED(uint8_t) PB_KEY_UNSIGNED(uint16_t) PB_KEY_UNSIGNED(uint32_t) PB_KEY_UNSIGNED(unsigned long long)
PB_KEY_SIGNED(int8_t) PB_KEY_SIGNED(int16_t) PB_KEY_SIGNED(int32_t) PB_KEY_SIGNED(long long)
template <> struct KeyOf<float> { static constexpr bool fp = true;
    __device__ static unsigned long long key(float v) { return key_of_f64((double)v); } };
template <> struct KeyOf<double> { static constexpr bool fp = true;
    __device__ static unsigned long long key(double v) { return key_of_f64(v); } };

constexpr unsigned long long KEY_MAX = 0xFFFFFFFFFFFFFFFFull, KEY_MIN = 0ull;

// block-level ladder: 2*NC keys per thread -> 2*NC global atomics per block
template <int NC>
__device__ void block_reduce_keys(unsigned long long (&kmin)[NC], unsigned long long (&kmax)[NC],
                                  unsigned long long* g_keys /* [NC mins][NC maxs] */) {
    __shared__ unsigned long long s_min[NC], s_max[NC];
    if (threadIdx.x < NC) { s_min[threadIdx.x] = KEY_MAX; s_max[threadIdx.x] = KEY_MIN; }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        unsigned long long a = kmin[c], b = kmax[c];
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long a2 = __shfl_xor_sync(0xffffffffu, a, o), b2 = __shfl_xor_sync(0xffffffffu, b, o);
            a = a2 < a ? a2 : a;
            b = b2 > b ? b2 : b;
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&s_min[c], a); atomicMax(&s_max[c], b); }
    }
    __syncthreads();
    if (threadIdx.x < NC) {
        atomicMin(&g_keys[threadIdx.x], s_min[threadIdx.x]);
        atomicMax(&g_keys[NC + threadIdx.x], s_max[threadIdx.x]);
    }
}

__global__ void init_keys_kernel(unsigned long long* keys, int nc) {
    if ((int)threadIdx.x < nc) keys[threadIdx.x] = KEY_MAX;
    else if ((int)threadIdx.x < 2 * nc) keys[threadIdx.x] = KEY_MIN;
}

// K5 fast path: packed Vec3f64 column (HashMapBuffer POSITION_3D), 16 B aligned, read as 16-byte chunks (two doubles).
// A warp step covers 96 consecutive chunks (64 points) with three fully coalesced loads per lane: chunk 96 j + 32 k + lane.
// The components held by a chunk depend on (k + 2 * lane) mod 3 only -- not on j -- so every lane keeps six private
// min/max pairs (one per load slot and half), updates them branch-free with strict compares (NaN never enters,
// bounds.rs:34-51) and maps them to x/y/z once at the end.  12 B-compares per 48 bytes: the kernel is HBM-bound.
__global__ void __launch_bounds__(256) bounds_flat_f64_kernel(const double2* __restrict__ data, unsigned long long n_doubles,
                                                              unsigned long long* g_keys) {
    double mnA[3], mxA[3], mnB[3], mxB[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { mnA[k] = mnB[k] = DBL_MAX; mxA[k] = mxB[k] = -DBL_MAX; }
    const unsigned long long n_chunks = n_doubles >> 1;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long n_steps = n_chunks / 96;
    auto upd = [&](int k, const double2 v) {
        mnA[k] = v.x < mnA[k] ? v.x : mnA[k]; mxA[k] = v.x > mxA[k] ? v.x : mxA[k];
        mnB[k] = v.y < mnB[k] ? v.y : mnB[k]; mxB[k] = v.y > mxB[k] ? v.y : mxB[k];
    };
    unsigned long long j = warp;
    for (; j + warps < n_steps; j += 2 * warps) {  // two steps (six loads) in flight
        const double2* p = data + j * 96 + lane;
        const double2* q = data + (j + warps) * 96 + lane;
        const double2 a0 = __ldg(p), a1 = __ldg(p + 32), a2 = __ldg(p + 64), b0 = __ldg(q), b1 = __ldg(q + 32), b2 = __ldg(q + 64);
        upd(0, a0); upd(1, a1); upd(2, a2);
        upd(0, b0); upd(1, b1); upd(2, b2);
    }
    for (; j < n_steps; j += warps) {
        const double2* p = data + j * 96 + lane;
        upd(0, __ldg(p)); upd(1, __ldg(p + 32)); upd(2, __ldg(p + 64));
    }
    // fold the slots into components: the first double of slot k is component (k + 2 * lane) mod 3, the second the next one
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned m = (k + 2u * lane) % 3u, m2 = (m + 1u) % 3u;
#pragma unroll
        for (unsigned c = 0; c < 3; ++c) {
            if (m == c) { if (mnA[k] < mn[c]) mn[c] = mnA[k]; if (mxA[k] > mx[c]) mx[c] = mxA[k]; }
            if (m2 == c) { if (mnB[k] < mn[c]) mn[c] = mnB[k]; if (mxB[k] > mx[c]) mx[c] = mxB[k]; }
        }
    }
    // tail: the chunks after the last whole step, one per thread of block 0, and an odd last double
    if (blockIdx.x == 0) {
        for (unsigned long long c = n_steps * 96 + threadIdx.x; c < n_chunks; c += blockDim.x) {
            const double2 v = __ldg(data + c);
            const unsigned m = (unsigned)((2 * c) % 3), m2 = (m + 1u) % 3u;
#pragma unroll
            for (unsigned cc = 0; cc < 3; ++cc) {
                if (m == cc) { if (v.x < mn[cc]) mn[cc] = v.x; if (v.x > mx[cc]) mx[cc] = v.x; }
                if (m2 == cc) { if (v.y < mn[cc]) mn[cc] = v.y; if (v.y > mx[cc]) mx[cc] = v.y; }
            }
        }
        if ((n_doubles & 1) && threadIdx.x == 0) {  // n_doubles = 3 * len may be odd
            const double a = reinterpret_cast<const double*>(data)[n_doubles - 1];
            const unsigned m = (unsigned)((n_doubles - 1) % 3);
#pragma unroll
            for (unsigned cc = 0; cc < 3; ++cc)
                if (m == cc) { if (a < mn[cc]) mn[cc] = a; if (a > mx[cc]) mx[cc] = a; }
        }
    }
    unsigned long long kmin[3], kmax[3];
    for (int k = 0; k < 3; ++k) { kmin[k] = key_of_f64(mn[k]); kmax[k] = key_of_f64(mx[k]); }
    block_reduce_keys<3>(kmin, kmax, g_keys);
}

template <class T>
__device__ __forceinline__ T ld_any(const uint8_t* p, bool aligned) {
    if (aligned) return *reinterpret_cast<const T*>(p);
    T v;
    uint8_t* b = reinterpret_cast<uint8_t*>(&v);
#pragma unroll
    for (int k = 0; k < (int)sizeof(T); ++k) b[k] = p[k];
    return v;
}

// K5/K6 generic: any scalar / Vec3 attribute at (base, stride); NaN never enters a float min/max
template <class T, int NC>
__global__ void __launch_bounds__(256) minmax_strided_kernel(const uint8_t* __restrict__ base, unsigned long long stride,
                                                             unsigned long long n, int aligned,
                                                             unsigned long long* g_keys) {
    unsigned long long kmin[NC], kmax[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { kmin[c] = KEY_MAX; kmax[c] = KEY_MIN; }
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const uint8_t* p = base + i * stride;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            T v = ld_any<T>(p + c * sizeof(T), aligned != 0);
            if (KeyOf<T>::fp && v != v) continue;
            unsigned long long k = KeyOf<T>::key(v);
            kmin[c] = k < kmin[c] ? k : kmin[c];
            kmax[c] = k > kmax[c] ? k : kmax[c];
        }
    }
    block_reduce_keys<NC>(kmin, kmax, g_keys);
}

template <int NC>
static void launch_minmax(uint32_t comp, const uint8_t* base, uint64_t stride, uint64_t n, int aligned,
                          unsigned long long* keys, int blocks, cudaStream_t st) {
    switch (comp) {
        case PB200_U8: minmax_strided_kernel<uint8_t, NC><<<blocks, 256, 0, st>>>(base, stride, n, aligned, keys); break;
        case PB200_I8: minmax_strided_kernel<int8_t, NC><<<blocks, 256, 0, st>>>(base, stride, n, aligned, keys); break;
        case PB200_U16: minmax_strided_kernel<uint16_t, NC><<<blocks, 256, 0, st>>>(base, stride, n, aligned, keys); break;
        case PB200_I16: minmax_strided_kernel<int16_t, NC><<<blocks, 256, 0, st>>>(base, stride, n, aligned, keys); break;
        case PB200_U32: minmax_strided_kernel<uint32_t, NC><<<blocks, 256, 0, st>>>(base, stride, n, aligned, keys); break;
        case PB200_I32: minmax_strided_kernel<int32_t, NC><<<blocks, 256, 0, st>>>(base, stride, n, aligned, keys); break;
        case PB200_U64: minmax_strided_kernel<unsigned long long, NC><<<blocks, 256, 0, st>>>(base, stride, n, aligned, keys); break;
        case PB200_I64: minmax_strided_kernel<long long, NC><<<blocks, 256, 0, st>>>(base, stride, n, aligned, keys); break;
        case PB200_F32: minmax_strided_kernel<float, NC><<<blocks, 256, 0, st>>>(base, stride, n, aligned, keys); break;
        default: minmax_strided_kernel<double, NC><<<blocks, 256, 0, st>>>(base, stride, n, aligned, keys); break;
    }
    g_launches++;
}

// Reduce attribute `idx` of `buf` (device or host memory) to 2*nc keys (host array `keys_out`).
static int reduce_attribute_keys(pb200_ctx* ctx, const pb200_buffer_desc* buf, int idx, unsigned long long* keys_out) {
    PB_DEVICE(ctx);
    const pb200_attr& a = buf->layout->attrs[(size_t)idx];
    const bool vec = is_cast_vec3(a.dtype);
    const uint32_t comp = vec ? vec3_component(a.dtype) : a.dtype;
    const int nc = vec ? 3 : 1;
    const uint64_t csize = pb200_dtype_size(comp, 0);
    void* scr = nullptr;
    PB_TRY(scratch(ctx, 4096, &scr));
    unsigned long long* d_keys = (unsigned long long*)scr + 64;
    init_keys_kernel<<<1, 32, 0, ctx->stream>>>(d_keys, nc);
    g_launches++;
    const uint64_t stride = buf->kind == PB200_INTERLEAVED ? buf->layout->size : a.size;
    const uint64_t off = buf->kind == PB200_INTERLEAVED ? a.offset : 0;
    const uint8_t* host_or_dev = buf->kind == PB200_INTERLEAVED ? (const uint8_t*)buf->aos : (const uint8_t*)buf->columns[idx];
    auto run = [&](const uint8_t* dbase, uint64_t n) {
        const uint8_t* p = dbase + off;
        const int aligned = (((uintptr_t)p % csize) == 0 && (stride % csize) == 0) ? 1 : 0;
        if (vec && comp == PB200_F64 && stride == 24 && ((uintptr_t)p & 15) == 0) {
            unsigned long long nd = 3ull * n;
            unsigned long long want = ((nd >> 1) + 256 * 4 - 1) / (256 * 4);
            int blocks = (int)(want < (unsigned long long)ctx->sm_count * 8 ? (want ? want : 1) : (unsigned long long)ctx->sm_count * 8);
            bounds_flat_f64_kernel<<<blocks, 256, 0, ctx->stream>>>((const double2*)p, nd, d_keys);
            g_launches++;
        } else {
            unsigned long long want = (n + 255) / 256;
            int blocks = (int)(want < (unsigned long long)ctx->sm_count * 8 ? (want ? want : 1) : (unsigned long long)ctx->sm_count * 8);
            if (nc == 3) launch_minmax<3>(comp, p, stride, n, aligned, d_keys, blocks, ctx->stream);
            else launch_minmax<1>(comp, p, stride, n, aligned, d_keys, blocks, ctx->stream);
        }
    };
    if (buf->memspace == PB200_DEVICE) {
        run(host_or_dev, buf->len);
    } else {  // stage the stream through device memory in 64 MB pieces
        uint64_t chunk = ((uint64_t)64 << 20) / (stride ? stride : 1);
        if (chunk == 0) chunk = 1;
        void* d_stage = nullptr;
        PB_CUDA(cudaMalloc(&d_stage, (size_t)(chunk * stride + 16)));
        for (uint64_t p0 = 0; p0 < buf->len; p0 += chunk) {
            uint64_t n = buf->len - p0 < chunk ? buf->len - p0 : chunk;
            cudaError_t e = cudaMemcpyAsync(d_stage, host_or_dev + p0 * stride, (size_t)(n * stride), cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) { cudaFree(d_stage); return cuda_error(e, "H2D staging"); }
            run((const uint8_t*)d_stage, n);
        }
        cudaStreamSynchronize(ctx->stream);
        cudaFree(d_stage);
    }
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaMemcpyAsync(ctx->h_scratch, d_keys, sizeof(unsigned long long) * 2 * nc, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(keys_out, ctx->h_scratch, sizeof(unsigned long long) * 2 * nc);
    return PB200_OK;
}

static int read_first_element(pb200_ctx* ctx, const pb200_buffer_desc* buf, int idx, uint8_t* out) {
    const pb200_attr& a = buf->layout->attrs[(size_t)idx];
    const uint8_t* p = buf->kind == PB200_INTERLEAVED ? (const uint8_t*)buf->aos + a.offset : (const uint8_t*)buf->columns[idx];
    if (buf->memspace == PB200_HOST) { memcpy(out, p, a.size); return PB200_OK; }
    PB_CUDA(cudaMemcpyAsync(ctx->h_scratch, p, a.size, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->h_scratch, a.size);
    return PB200_OK;
}

__global__ void __launch_bounds__(256) morton_kernel(const uint8_t* __restrict__ base, unsigned long long stride,
                                                     unsigned long long n, double bx, double by, double bz, double sx,
                                                     double sy, double sz, unsigned long long* __restrict__ out) {
    auto expand = [](unsigned long long val) {  // math/bitmanip.rs:2-10
        val &= 0x1FFFFFull;
        val = (val | (val << 32)) & 0x00FF00000000FFFFull;
        val = (val | (val << 16)) & 0x00FF0000FF0000FFull;
        val = (val | (val << 8)) & 0xF00F00F00F00F00Full;
        val = (val | (val << 4)) & 0x30C30C30C30C30C3ull;
        val = (val | (val << 2)) & 0x1249249249249249ull;
        return val;
    };
    auto quant = [](double p, double b, double s) {
        double t = __dmul_rn(__dsub_rn(p, b), s);
        if (!(t > 0.0)) return 0ull;
        unsigned long long q = t >= 2097151.0 ? 2097151ull : __double2ull_rz(t);
        return q;
    };
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const double* p = reinterpret_cast<const double*>(base + i * stride);
        out[i] = (expand(quant(p[0], bx, sx)) << 2) | (expand(quant(p[1], by, sy)) << 1) | expand(quant(p[2], bz, sz));
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_calculate_bounds(pb200_ctx* ctx, const pb200_buffer_desc* buf, double out_min[3], double out_max[3],
                           int* is_some) {
    if (!ctx || !out_min || !out_max || !is_some) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(buf, "buffer"));
    *is_some = 0;
    if (buf->len == 0) return PB200_OK;  // bounds.rs:12-14
    int pi = pb200_layout_index_by_name(buf->layout, "Position3D");
    if (pi < 0) return PB200_OK;         // bounds.rs:15-21
    const uint32_t d = buf->layout->attrs[(size_t)pi].dtype;
    if (!is_cast_vec3(d))                // get_generic_converter panics (attribute_conversion.rs:267-269)
        return set_error(PB200_ERR_NO_CONVERSION, "Invalid conversion dtype %u -> Vec3<f64>", d);
    unsigned long long keys[6];
    PB_TRY(reduce_attribute_keys(ctx, buf, pi, keys));
    for (int c = 0; c < 3; ++c) {
        // integer and f32 components were widened exactly (value-preserving `as f64`, bounds.rs:62-65)
        const uint32_t comp = vec3_component(d);
        double mn, mx;
        if (comp == PB200_F64 || comp == PB200_F32) {
            mn = keys[c] == KEY_MAX ? DBL_MAX : f64_of_key(keys[c]);
            mx = keys[3 + c] == KEY_MIN ? -DBL_MAX : f64_of_key(keys[3 + c]);
        } else if (comp == PB200_I32) {
            mn = (double)(long long)(keys[c] ^ 0x8000000000000000ull);
            mx = (double)(long long)(keys[3 + c] ^ 0x8000000000000000ull);
        } else {
            mn = (double)keys[c];
            mx = (double)keys[3 + c];
        }
        out_min[c] = mn;
        out_max[c] = mx;
    }
    if (out_min[0] > out_max[0] || out_min[1] > out_max[1] || out_min[2] > out_max[2])  // math/bounds.rs:21-26
        return set_error(PB200_ERR_INVALID, "AABB::from_min_max: Minimum position must be <= maximum position!");
    *is_some = 1;
    return PB200_OK;
}

static int minmax_attribute_impl(pb200_ctx* ctx, const pb200_buffer_desc* buf, const char* name, uint32_t dtype, bool seed_rule,
                                 void* out_min, void* out_max, int* is_some);

int pb200_minmax_attribute(pb200_ctx* ctx, const pb200_buffer_desc* buf, const char* name, uint32_t dtype,
                           void* out_min, void* out_max, int* is_some) {
    return minmax_attribute_impl(ctx, buf, name, dtype, true, out_min, out_max, is_some);
}

int pb200_minmax_attribute_partial(pb200_ctx* ctx, const pb200_buffer_desc* buf, const char* name, uint32_t dtype,
                                   void* out_min, void* out_max, int* is_some) {
    return minmax_attribute_impl(ctx, buf, name, dtype, false, out_min, out_max, is_some);
}

static int minmax_attribute_impl(pb200_ctx* ctx, const pb200_buffer_desc* buf, const char* name, uint32_t dtype, bool seed_rule,
                                 void* out_min, void* out_max, int* is_some) {
    if (!ctx || !name || !out_min || !out_max || !is_some) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(buf, "buffer"));
    *is_some = 0;
    if (pb200_layout_index_by_name(buf->layout, name) < 0)  // minmax.rs:17-26
        return set_error(PB200_ERR_ATTR_NOT_FOUND, "Attribute %s not contained in PointLayout of buffer", name);
    int idx = pb200_layout_index_of(buf->layout, name, dtype);
    if (idx < 0)  // view_attribute::<T> panics on a datatype mismatch
        return set_error(PB200_ERR_ATTR_NOT_FOUND, "Attribute %s with dtype %u not contained in PointLayout of buffer", name, dtype);
    if (!is_scalar(dtype) && !is_cast_vec3(dtype))
        return set_error(PB200_ERR_UNSUPPORTED, "dtype %u has no MinMax implementation", dtype);
    if (buf->len == 0) return PB200_OK;
    const bool vec = is_cast_vec3(dtype);
    const uint32_t comp = vec ? vec3_component(dtype) : dtype;
    const int nc = vec ? 3 : 1;
    const size_t cs = (size_t)pb200_dtype_size(comp, 0);
    unsigned long long keys[6];
    PB_TRY(reduce_attribute_keys(ctx, buf, idx, keys));
    uint8_t first[32];
    memset(first, 0, sizeof first);  // (0.0 is not NaN: without the seed rule nothing below replaces a result)
    if (seed_rule) PB_TRY(read_first_element(ctx, buf, idx, first));
    for (int c = 0; c < nc; ++c) {
        uint8_t* omn = (uint8_t*)out_min + c * cs;
        uint8_t* omx = (uint8_t*)out_max + c * cs;
        const unsigned long long kmn = keys[c], kmx = keys[nc + c];
        switch (comp) {
            case PB200_F64: {
                double f; memcpy(&f, first + c * cs, 8);
                // the first value seeds (min,max); a NaN seed is never replaced (math/minmax.rs:62-96)
                double mn = (f != f) ? f : f64_of_key(kmn), mx = (f != f) ? f : f64_of_key(kmx);
                memcpy(omn, &mn, 8); memcpy(omx, &mx, 8);
                break;
            }
            case PB200_F32: {
                float f; memcpy(&f, first + c * cs, 4);
                float mn = (f != f) ? f : (float)f64_of_key(kmn), mx = (f != f) ? f : (float)f64_of_key(kmx);
                memcpy(omn, &mn, 4); memcpy(omx, &mx, 4);
                break;
            }
            case PB200_I8: case PB200_I16: case PB200_I32: case PB200_I64: {
                long long mn = (long long)(kmn ^ 0x8000000000000000ull), mx = (long long)(kmx ^ 0x8000000000000000ull);
                memcpy(omn, &mn, cs); memcpy(omx, &mx, cs);  // little endian: low bytes
                break;
            }
            default: {
                memcpy(omn, &kmn, cs); memcpy(omx, &kmx, cs);
                break;
            }
        }
    }
    *is_some = 1;
    return PB200_OK;
}

int pb200_morton_codes(pb200_ctx* ctx, const pb200_buffer_desc* buf, const double bmin[3], const double bmax[3],
                       uint64_t* codes_out) {
    if (!ctx || !bmin || !bmax || !codes_out) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(buf, "buffer"));
    PB_DEVICE(ctx);
    int pi = pb200_layout_index_of(buf->layout, "Position3D", PB200_VEC3F64);
    if (pi < 0) return set_error(PB200_ERR_ATTR_NOT_FOUND, "buffer has no Vec3f64 Position3D attribute");
    if (buf->memspace != PB200_DEVICE) return set_error(PB200_ERR_UNSUPPORTED, "pb200_morton_codes needs device memory");
    if (buf->len == 0) return PB200_OK;
    const pb200_attr& a = buf->layout->attrs[(size_t)pi];
    const uint64_t stride = buf->kind == PB200_INTERLEAVED ? buf->layout->size : a.size;
    const uint8_t* p = buf->kind == PB200_INTERLEAVED ? (const uint8_t*)buf->aos + a.offset : (const uint8_t*)buf->columns[pi];
    if (((uintptr_t)p & 7) || (stride & 7)) return set_error(PB200_ERR_UNSUPPORTED, "positions must be 8-byte aligned");
    double s[3];
    for (int c = 0; c < 3; ++c) {
        double e = bmax[c] - bmin[c];
        s[c] = e > 0.0 ? 2097152.0 / e : 0.0;
    }
    unsigned long long want = (buf->len + 255) / 256;
    int blocks = (int)(want < (unsigned long long)ctx->sm_count * 8 ? want : (unsigned long long)ctx->sm_count * 8);
    morton_kernel<<<blocks, 256, 0, ctx->stream>>>(p, stride, buf->len, bmin[0], bmin[1], bmin[2], s[0], s[1], s[2],
                                                   (unsigned long long*)codes_out);
    g_launches++;
    PB_CUDA(cudaGetLastError());
    return PB200_OK;
}

}  // extern "C"
