// ransac.cu -- inlier counting of pasture-algorithms/src/segmentation.rs (ransac_plane_* / ransac_line_*).
//
// The reference evaluates its num_of_iterations hypotheses one after the other (or one per rayon task), each one
// streaming the whole POSITION_3D attribute (:98-138).  Here all hypotheses of a batch live in shared memory and the
// positions are read from HBM ONCE per batch: every thread keeps 4 points in registers and walks the hypotheses,
// warps count with redux.sync into private shared-memory rows, blocks flush with one 64-bit atomic per hypothesis.
//
// Parity: the decision is `distance < threshold` on the f64 value the reference computes (:31-44: a division for
// planes, sqrt + division for lines).  The kernels evaluate the numerator exactly as written (no FMA), then classify
// with a guard band: a point whose numerator is outside [t(1-8eps), t(1+8eps)], t = threshold * denominator, is
// decided without the division (rounding cannot move it across the threshold); only points inside the band pay
// the exact sqrt/division.  The result is bit-identical to evaluating the reference expression everywhere.
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <cmath>

#include "internal.h"

namespace pb200 {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_PPT = 4;      // points per thread
constexpr int RS_BATCH = 256;  // hypotheses per launch

using RBuf = DevTmp;  // stream-ordered temporaries, recycled between calls

// one hypothesis, pre-digested on the host. plane: m = a,b,c,d ; line: m = first xyz, v = second - first
struct Hyp {
    double m[6];
    double den;       // e = sqrt(a^2+b^2+c^2)  |  |second - first|
    double lo, hi;    // guard band on the numerator (plane) / squared numerator (line); lo > hi = always exact
    double thr;
};

template <int KIND>
__device__ __forceinline__ bool exact_inlier(const Hyp& h, double x, double y, double z) {
    if (KIND == 0) {  // distance_point_plane, segmentation.rs:31-35
        const double d = fabs(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(h.m[0], x), __dmul_rn(h.m[1], y)), __dmul_rn(h.m[2], z)), h.m[3]));
        return __ddiv_rn(d, h.den) < h.thr;
    } else {  // distance_point_line, :39-44
        const double wx = __dsub_rn(h.m[0], x), wy = __dsub_rn(h.m[1], y), wz = __dsub_rn(h.m[2], z);
        const double cx = __dsub_rn(__dmul_rn(h.m[4], wz), __dmul_rn(h.m[5], wy));
        const double cy = __dsub_rn(__dmul_rn(h.m[5], wx), __dmul_rn(h.m[3], wz));
        const double cz = __dsub_rn(__dmul_rn(h.m[3], wy), __dmul_rn(h.m[4], wx));
        const double s = __dadd_rn(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)), __dmul_rn(cz, cz));
        return __ddiv_rn(__dsqrt_rn(s), h.den) < h.thr;
    }
}

template <int KIND>
__device__ __forceinline__ bool inlier(const Hyp& h, double x, double y, double z) {
    double q;
    if (KIND == 0) {
        q = fabs(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(h.m[0], x), __dmul_rn(h.m[1], y)), __dmul_rn(h.m[2], z)), h.m[3]));
    } else {
        const double wx = __dsub_rn(h.m[0], x), wy = __dsub_rn(h.m[1], y), wz = __dsub_rn(h.m[2], z);
        const double cx = __dsub_rn(__dmul_rn(h.m[4], wz), __dmul_rn(h.m[5], wy));
        const double cy = __dsub_rn(__dmul_rn(h.m[5], wx), __dmul_rn(h.m[3], wz));
        const double cz = __dsub_rn(__dmul_rn(h.m[3], wy), __dmul_rn(h.m[4], wx));
        q = __dadd_rn(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)), __dmul_rn(cz, cz));
    }
    if (q < h.lo) return true;
    if (q > h.hi) return false;
    return exact_inlier<KIND>(h, x, y, z);  // inside the band, NaN, or a degenerate model
}

template <int KIND>
__global__ void __launch_bounds__(RS_THREADS) ransac_rank_kernel(const uint8_t* __restrict__ base, unsigned long long stride,
                                                                 unsigned long long n, const Hyp* __restrict__ hyps, int n_hyp,
                                                                 unsigned long long* __restrict__ rankings) {
    __shared__ Hyp s_h[RS_BATCH];
    __shared__ uint32_t s_cnt[RS_WARPS][RS_BATCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < n_hyp; i += RS_THREADS) s_h[i] = hyps[i];
    for (int i = threadIdx.x; i < RS_WARPS * RS_BATCH; i += RS_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const unsigned long long chunk = (unsigned long long)RS_THREADS * RS_PPT;
    for (unsigned long long c0 = (unsigned long long)blockIdx.x * chunk; c0 < n; c0 += (unsigned long long)gridDim.x * chunk) {
        double px[RS_PPT], py[RS_PPT], pz[RS_PPT];
        bool ok[RS_PPT];
#pragma unroll
        for (int j = 0; j < RS_PPT; ++j) {
            const unsigned long long i = c0 + (unsigned long long)j * RS_THREADS + threadIdx.x;
            ok[j] = i < n;
            const double* p = reinterpret_cast<const double*>(base + (ok[j] ? i : 0) * stride);
            px[j] = p[0]; py[j] = p[1]; pz[j] = p[2];
        }
        for (int h = 0; h < n_hyp; ++h) {
            const Hyp& H = s_h[h];
            unsigned c = 0;
#pragma unroll
            for (int j = 0; j < RS_PPT; ++j) c += (ok[j] && inlier<KIND>(H, px[j], py[j], pz[j])) ? 1u : 0u;
            c = __reduce_add_sync(0xFFFFFFFFu, c);
            if (lane == 0) s_cnt[warp][h] += c;
        }
    }
    __syncthreads();
    for (int h = threadIdx.x; h < n_hyp; h += RS_THREADS) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) t += s_cnt[w][h];
        if (t) atomicAdd(&rankings[h], t);
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) ransac_flags_kernel(const uint8_t* __restrict__ base, unsigned long long stride, unsigned long long n,
                                                           Hyp h, uint8_t* __restrict__ flags) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const double* p = reinterpret_cast<const double*>(base + i * stride);
        flags[i] = inlier<KIND>(h, p[0], p[1], p[2]) ? 1 : 0;
    }
}

__global__ void gather_samples_kernel(const uint8_t* __restrict__ base, unsigned long long stride, const unsigned long long* __restrict__ idx,
                                      uint32_t count, double* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const double* p = reinterpret_cast<const double*>(base + idx[i] * stride);
        out[3 * i] = p[0]; out[3 * i + 1] = p[1]; out[3 * i + 2] = p[2];
    }
}

static Hyp make_hyp(int kind, const double* m, double thr) {
    Hyp h{};
    h.thr = thr;
    volatile double den2;  // keep every product and sum a separately rounded f64 operation
    if (kind == 0) {
        for (int c = 0; c < 4; ++c) h.m[c] = m[c];
        volatile double aa = m[0] * m[0], bb = m[1] * m[1], cc = m[2] * m[2];
        volatile double s = aa + bb;
        den2 = s + cc;
    } else {
        for (int c = 0; c < 3; ++c) { h.m[c] = m[c]; volatile double v = m[3 + c] - m[c]; h.m[3 + c] = v; }
        volatile double aa = h.m[3] * h.m[3], bb = h.m[4] * h.m[4], cc = h.m[5] * h.m[5];
        volatile double s = aa + bb;
        den2 = s + cc;
    }
    h.den = std::sqrt((double)den2);
    // guard band. |q_computed/den - fl(q_computed/den)| <= eps/2 relative, t = thr*den carries eps/2 more, and for lines
    // the comparison is on the square (sqrt halves relative distances): 8 eps on t (32 eps on t^2) is far outside
    // anything rounding can do, far inside anything that costs measurable time.
    const double eps = 2.220446049250313e-16;
    const double t = thr * h.den;
    h.lo = -INFINITY; h.hi = INFINITY;  // nothing is outside this band: always evaluate the reference expression
    if (thr > 1e-290 && std::isfinite(t) && t > 1e-140 && t < 1e140) {
        if (kind == 0) { h.lo = t * (1.0 - 8 * eps); h.hi = t * (1.0 + 8 * eps); }
        else { h.lo = t * t * (1.0 - 32 * eps); h.hi = t * t * (1.0 + 32 * eps); }
    }
    return h;
}

struct Positions {
    RBuf staged;
    const uint8_t* base = nullptr;
    uint64_t stride = 0;
};

static int ransac_positions(pb200_ctx* ctx, const pb200_buffer_desc* buf, Positions* P) {
    const int pi = pb200_layout_index_of(buf->layout, "Position3D", PB200_VEC3F64);
    if (pi < 0) return set_error(PB200_ERR_ATTR_NOT_FOUND, "buffer has no Vec3f64 Position3D attribute (view_attribute::<Vector3<f64>> would panic)");
    const pb200_attr& at = buf->layout->attrs[(size_t)pi];
    const uint8_t* p;
    if (buf->kind == PB200_INTERLEAVED) { P->stride = buf->layout->size; p = (const uint8_t*)buf->aos; }
    else { P->stride = at.size; p = (const uint8_t*)buf->columns[pi]; }
    if (buf->memspace == PB200_HOST) {
        const size_t bytes = (size_t)(buf->len * P->stride);
        PB_CUDA(P->staged.alloc(ctx->stream, bytes));
        PB_CUDA(cudaMemcpyAsync(P->staged.p, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
        p = (const uint8_t*)P->staged.p;
    }
    P->base = p + (buf->kind == PB200_INTERLEAVED ? at.offset : 0);
    if (((uintptr_t)P->base & 7) || (P->stride & 7)) return set_error(PB200_ERR_UNSUPPORTED, "POSITION_3D must be 8-byte aligned in memory");
    return PB200_OK;
}

static int check_kind_len(int kind, const pb200_buffer_desc* buf) {
    if (kind != PB200_RANSAC_PLANE && kind != PB200_RANSAC_LINE) return set_error(PB200_ERR_INVALID, "kind must be PB200_RANSAC_PLANE or PB200_RANSAC_LINE");
    if (kind == PB200_RANSAC_PLANE && buf->len < 3)  // segmentation.rs:184-186
        return set_error(PB200_ERR_TOO_FEW_POINTS, "buffer needs to include at least 3 points to generate a plane.");
    if (kind == PB200_RANSAC_LINE && buf->len < 2)  // :295-297
        return set_error(PB200_ERR_TOO_FEW_POINTS, "buffer needs to include at least 2 points to generate a line.");
    return PB200_OK;
}

static int rank_models(pb200_ctx* ctx, const Positions& P, uint64_t n, int kind, const double* models, uint64_t n_models,
                       double thr, uint64_t* rankings) {
    const int w = kind == 0 ? 4 : 6;
    std::vector<Hyp> hyps((size_t)n_models);
    for (uint64_t h = 0; h < n_models; ++h) hyps[(size_t)h] = make_hyp(kind, models + w * h, thr);
    RBuf d_h, d_r;
    PB_CUDA(d_h.alloc(ctx->stream, sizeof(Hyp) * (size_t)n_models));
    PB_CUDA(d_r.alloc(ctx->stream, 8 * (size_t)n_models));
    PB_CUDA(cudaMemcpyAsync(d_h.p, hyps.data(), sizeof(Hyp) * (size_t)n_models, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(cudaMemsetAsync(d_r.p, 0, 8 * (size_t)n_models, ctx->stream));
    const unsigned long long chunk = (unsigned long long)RS_THREADS * RS_PPT;
    const unsigned long long want = (n + chunk - 1) / chunk, cap = (unsigned long long)ctx->sm_count * 4;
    const unsigned blocks = (unsigned)(want < cap ? want : cap);
    for (uint64_t h0 = 0; h0 < n_models; h0 += RS_BATCH) {
        const int nh = (int)((n_models - h0) < RS_BATCH ? (n_models - h0) : RS_BATCH);
        if (kind == 0) ransac_rank_kernel<0><<<blocks, RS_THREADS, 0, ctx->stream>>>(P.base, P.stride, n, (const Hyp*)d_h.p + h0, nh, (unsigned long long*)d_r.p + h0);
        else ransac_rank_kernel<1><<<blocks, RS_THREADS, 0, ctx->stream>>>(P.base, P.stride, n, (const Hyp*)d_h.p + h0, nh, (unsigned long long*)d_r.p + h0);
        g_launches++;
    }
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaMemcpyAsync(rankings, d_r.p, 8 * (size_t)n_models, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
}

static int model_inliers(pb200_ctx* ctx, const Positions& P, uint64_t n, int kind, const double* model, double thr, bool host,
                         uint64_t* indices, uint64_t capacity, uint64_t* count) {
    const Hyp h = make_hyp(kind, model, thr);
    RBuf d_flags, d_idx, d_num, d_tmp;
    PB_CUDA(d_flags.alloc(ctx->stream, (size_t)n));
    PB_CUDA(d_num.alloc(ctx->stream, 8));
    const unsigned long long cap = (unsigned long long)ctx->sm_count * 16;
    const unsigned blocks = (unsigned)(((n + 255) / 256) < cap ? ((n + 255) / 256) : cap);
    if (kind == 0) ransac_flags_kernel<0><<<blocks, 256, 0, ctx->stream>>>(P.base, P.stride, n, h, (uint8_t*)d_flags.p);
    else ransac_flags_kernel<1><<<blocks, 256, 0, ctx->stream>>>(P.base, P.stride, n, h, (uint8_t*)d_flags.p);
    g_launches++;
    PB_CUDA(cudaGetLastError());
    // the select writes up to n indices: go through a device buffer of full size unless the caller's device buffer has room
    uint64_t* d_out = indices;
    if (host || capacity < n) {
        PB_CUDA(d_idx.alloc(ctx->stream, 8 * (size_t)n));
        d_out = (uint64_t*)d_idx.p;
    }
    thrust::counting_iterator<unsigned long long> it(0);
    size_t tmp = 0;
    cub::DeviceSelect::Flagged(nullptr, tmp, it, (const uint8_t*)d_flags.p, (unsigned long long*)d_out, (unsigned long long*)d_num.p, (long long)n, ctx->stream);
    PB_CUDA(d_tmp.alloc(ctx->stream, tmp));
    PB_CUDA(cub::DeviceSelect::Flagged(d_tmp.p, tmp, it, (const uint8_t*)d_flags.p, (unsigned long long*)d_out, (unsigned long long*)d_num.p, (long long)n, ctx->stream));
    g_launches++;
    unsigned long long m = 0;
    PB_CUDA(cudaMemcpyAsync(&m, d_num.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    *count = m;
    if (m > capacity) return set_error(PB200_ERR_RANGE, "indices capacity %llu is smaller than the number of inliers %llu", (unsigned long long)capacity, m);
    if (d_out != indices && m)
        PB_CUDA(cudaMemcpyAsync(indices, d_out, 8 * (size_t)m, host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
}

static void model_from_samples(int kind, const double* p, double* m) {
    if (kind == 0) {  // generate_rng_plane, segmentation.rs:60-76
        const double *pa = p, *pb = p + 3, *pc = p + 6;
        volatile double v1[3], v2[3];
        for (int c = 0; c < 3; ++c) { v1[c] = pb[c] - pa[c]; v2[c] = pc[c] - pa[c]; }
        volatile double t0, t1;
        double nrm[3];
        t0 = v1[1] * v2[2]; t1 = v1[2] * v2[1]; nrm[0] = t0 - t1;
        t0 = v1[2] * v2[0]; t1 = v1[0] * v2[2]; nrm[1] = t0 - t1;
        t0 = v1[0] * v2[1]; t1 = v1[1] * v2[0]; nrm[2] = t0 - t1;
        volatile double a = nrm[0] * pa[0], b = nrm[1] * pa[1], c = nrm[2] * pa[2];
        volatile double s = a + b;
        volatile double dot = s + c;
        m[0] = nrm[0]; m[1] = nrm[1]; m[2] = nrm[2]; m[3] = -dot;
    } else {  // generate_rng_line, :91-95
        for (int c = 0; c < 6; ++c) m[c] = p[c];
    }
}

static uint64_t splitmix64(uint64_t seed, uint64_t j) {
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

static int models_from_sample_indices(pb200_ctx* ctx, const Positions& P, int kind, const uint64_t* samples, uint64_t n_models, double* models) {
    const int per = kind == 0 ? 3 : 2, w = kind == 0 ? 4 : 6;
    const size_t cnt = (size_t)n_models * per;
    RBuf d_s, d_p;
    PB_CUDA(d_s.alloc(ctx->stream, 8 * cnt));
    PB_CUDA(d_p.alloc(ctx->stream, 24 * cnt));
    PB_CUDA(cudaMemcpyAsync(d_s.p, samples, 8 * cnt, cudaMemcpyHostToDevice, ctx->stream));
    gather_samples_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->stream>>>(P.base, P.stride, (const unsigned long long*)d_s.p, (uint32_t)cnt, (double*)d_p.p);
    g_launches++;
    std::vector<double> pts(3 * cnt);
    PB_CUDA(cudaMemcpyAsync(pts.data(), d_p.p, 24 * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint64_t h = 0; h < n_models; ++h) model_from_samples(kind, pts.data() + 3 * per * h, models + w * h);
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_ransac_rank_samples(pb200_ctx* ctx, const pb200_buffer_desc* buf, int kind, const uint64_t* samples, uint64_t n_models,
                              double distance_threshold, double* models_out, uint64_t* rankings_out) {
    if (!ctx || !samples || !models_out || !rankings_out) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(buf, "buffer"));
    PB_TRY(check_kind_len(kind, buf));
    if (n_models == 0) return PB200_OK;
    if (n_models > (1u << 24)) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^24 models per call");
    const int per = kind == 0 ? 3 : 2;
    for (uint64_t i = 0; i < n_models * per; ++i)
        if (samples[i] >= buf->len) return set_error(PB200_ERR_RANGE, "sample index %llu out of bounds", (unsigned long long)samples[i]);
    PB_DEVICE(ctx);
    Positions P;
    PB_TRY(ransac_positions(ctx, buf, &P));
    PB_TRY(models_from_sample_indices(ctx, P, kind, samples, n_models, models_out));
    return rank_models(ctx, P, buf->len, kind, models_out, n_models, distance_threshold, rankings_out);
}

int pb200_ransac_rank_models(pb200_ctx* ctx, const pb200_buffer_desc* buf, int kind, const double* models, uint64_t n_models,
                             double distance_threshold, uint64_t* rankings_out) {
    if (!ctx || !models || !rankings_out) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(buf, "buffer"));
    if (kind != PB200_RANSAC_PLANE && kind != PB200_RANSAC_LINE) return set_error(PB200_ERR_INVALID, "bad kind");
    if (n_models == 0) return PB200_OK;
    if (n_models > (1u << 24)) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^24 models per call");
    if (buf->len == 0) { memset(rankings_out, 0, 8 * (size_t)n_models); return PB200_OK; }
    PB_DEVICE(ctx);
    Positions P;
    PB_TRY(ransac_positions(ctx, buf, &P));
    return rank_models(ctx, P, buf->len, kind, models, n_models, distance_threshold, rankings_out);
}

int pb200_ransac_inliers(pb200_ctx* ctx, const pb200_buffer_desc* buf, int kind, const double* model, double distance_threshold,
                         uint64_t* indices_out, uint64_t capacity, uint64_t* num_inliers) {
    if (!ctx || !model || !num_inliers || (!indices_out && capacity)) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(buf, "buffer"));
    if (kind != PB200_RANSAC_PLANE && kind != PB200_RANSAC_LINE) return set_error(PB200_ERR_INVALID, "bad kind");
    *num_inliers = 0;
    if (buf->len == 0) return PB200_OK;
    PB_DEVICE(ctx);
    Positions P;
    PB_TRY(ransac_positions(ctx, buf, &P));
    return model_inliers(ctx, P, buf->len, kind, model, distance_threshold, buf->memspace == PB200_HOST, indices_out, capacity, num_inliers);
}

int pb200_ransac(pb200_ctx* ctx, const pb200_buffer_desc* buf, int kind, double distance_threshold, uint64_t num_of_iterations,
                 uint64_t seed, double* model_out, uint64_t* ranking_out, uint64_t* indices_out, uint64_t capacity) {
    if (!ctx || !model_out || !ranking_out) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(buf, "buffer"));
    PB_TRY(check_kind_len(kind, buf));
    if (num_of_iterations == 0) return set_error(PB200_ERR_INVALID, "num_of_iterations must be > 0 (max_by(..).unwrap() on an empty iterator panics)");
    if (num_of_iterations > (1u << 24)) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^24 iterations per call");
    PB_DEVICE(ctx);
    const uint64_t n = buf->len;
    const int per = kind == 0 ? 3 : 2, w = kind == 0 ? 4 : 6;
    std::vector<uint64_t> samples((size_t)num_of_iterations * per);
    uint64_t j = 0;
    for (uint64_t h = 0; h < num_of_iterations; ++h) {  // the reference's redraw loops, :50-59 / :83-89
        uint64_t r1 = splitmix64(seed, j++) % n, r2 = splitmix64(seed, j++) % n;
        while (r1 == r2) r2 = splitmix64(seed, j++) % n;
        samples[(size_t)h * per] = r1; samples[(size_t)h * per + 1] = r2;
        if (kind == 0) {
            uint64_t r3 = splitmix64(seed, j++) % n;
            while (r2 == r3 || r1 == r3) r3 = splitmix64(seed, j++) % n;
            samples[(size_t)h * per + 2] = r3;
        }
    }
    Positions P;
    PB_TRY(ransac_positions(ctx, buf, &P));
    std::vector<double> models((size_t)num_of_iterations * w);
    std::vector<uint64_t> ranks((size_t)num_of_iterations);
    PB_TRY(models_from_sample_indices(ctx, P, kind, samples.data(), num_of_iterations, models.data()));
    PB_TRY(rank_models(ctx, P, n, kind, models.data(), num_of_iterations, distance_threshold, ranks.data()));
    size_t best = 0;
    for (size_t h = 0; h < ranks.size(); ++h)
        if (ranks[h] >= ranks[best]) best = h;  // Iterator::max_by keeps the last of equal maxima
    for (int c = 0; c < w; ++c) model_out[c] = models[best * w + c];
    *ranking_out = ranks[best];
    if (!indices_out) return PB200_OK;
    uint64_t cnt = 0;
    return model_inliers(ctx, P, n, kind, model_out, distance_threshold, buf->memspace == PB200_HOST, indices_out, capacity, &cnt);
}

}  // extern "C"
