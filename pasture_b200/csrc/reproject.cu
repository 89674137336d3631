// reproject.cu -- per-point coordinate transform (pasture-algorithms/src/reprojection.rs:38-45,132-146,201-227).
// The reference calls PROJ (proj_create_crs_to_crs + proj_trans) per point; PROJ strings cannot run on the GPU, so
// the boundary takes an enumerated operation pipeline. The pipeline for the one CRS pair the reference's tests pin
// (EPSG:4326 -> EPSG:3309, reprojection.rs:275-289) is WGS84 geodetic -> ECEF -> NAD27 shift -> Clarke 1866
// geodetic -> Albers equal-area; z passes through unchanged.  HBM: 24 B in + 24 B out per point.
#include "internal.h"

namespace pb200 {

constexpr int MAX_PROJ_OPS = 8;
struct ProjPipeline { uint32_t n; pb200_proj_op ops[MAX_PROJ_OPS]; };

#define PB_PI 3.14159265358979323846

__device__ __forceinline__ double albers_q(double e, double sinphi) {  // Snyder 3-12
    const double es = e * sinphi;
    return (1.0 - e * e) * (sinphi / (1.0 - es * es) - (1.0 / (2.0 * e)) * log((1.0 - es) / (1.0 + es)));
}
__device__ __forceinline__ double tmerc_M(double a, double e2, double phi) {  // Snyder 3-21
    const double e4 = e2 * e2, e6 = e4 * e2;
    return a * ((1.0 - e2 / 4.0 - 3.0 * e4 / 64.0 - 5.0 * e6 / 256.0) * phi -
                (3.0 * e2 / 8.0 + 3.0 * e4 / 32.0 + 45.0 * e6 / 1024.0) * sin(2.0 * phi) +
                (15.0 * e4 / 256.0 + 45.0 * e6 / 1024.0) * sin(4.0 * phi) - (35.0 * e6 / 3072.0) * sin(6.0 * phi));
}

__global__ void __launch_bounds__(256) reproject_kernel(const uint8_t* __restrict__ src, unsigned long long sstride,
                                                        uint8_t* __restrict__ dst, unsigned long long dstride,
                                                        unsigned long long n, const __grid_constant__ ProjPipeline pl) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const double* in = reinterpret_cast<const double*>(src + i * sstride);
        double v0 = in[0], v1 = in[1], v2 = in[2];
        const double z_in = v2;
        for (uint32_t k = 0; k < pl.n; ++k) {
            const double* p = pl.ops[k].p;
            switch (pl.ops[k].kind) {
                case PB200_PROJ_AFFINE: {
                    const double r0 = p[0] * v0 + p[1] * v1 + p[2] * v2 + p[9];
                    const double r1 = p[3] * v0 + p[4] * v1 + p[5] * v2 + p[10];
                    const double r2 = p[6] * v0 + p[7] * v1 + p[8] * v2 + p[11];
                    v0 = r0; v1 = r1; v2 = r2;
                    break;
                }
                case PB200_PROJ_GEODETIC_TO_ECEF: {
                    const double a = p[0], f = 1.0 / p[1], e2 = f * (2.0 - f);
                    const double lat = v0 * (PB_PI / 180.0), lon = v1 * (PB_PI / 180.0), h = v2;
                    const double sl = sin(lat), cl = cos(lat);
                    const double N = a / sqrt(1.0 - e2 * sl * sl);
                    v0 = (N + h) * cl * cos(lon);
                    v1 = (N + h) * cl * sin(lon);
                    v2 = (N * (1.0 - e2) + h) * sl;
                    break;
                }
                case PB200_PROJ_ECEF_TO_GEODETIC: {
                    const double a = p[0], f = 1.0 / p[1], e2 = f * (2.0 - f);
                    const double X = v0, Y = v1, Z = v2;
                    const double lon = atan2(Y, X);
                    const double pr = sqrt(X * X + Y * Y);
                    double lat = atan2(Z, pr * (1.0 - e2));
                    double h = 0.0;
                    for (int it = 0; it < 8; ++it) {
                        const double sl = sin(lat);
                        const double N = a / sqrt(1.0 - e2 * sl * sl);
                        h = pr / cos(lat) - N;
                        lat = atan2(Z, pr * (1.0 - e2 * N / (N + h)));
                    }
                    v0 = lat; v1 = lon; v2 = h;
                    break;
                }
                case PB200_PROJ_ALBERS_FWD: {
                    const double a = p[0], f = 1.0 / p[1], e2 = f * (2.0 - f), e = sqrt(e2);
                    const double phi1 = p[2], phi2 = p[3], phi0 = p[4], lam0 = p[5], x0 = p[6], y0 = p[7];
                    const double m1 = cos(phi1) / sqrt(1.0 - e2 * sin(phi1) * sin(phi1));
                    const double m2 = cos(phi2) / sqrt(1.0 - e2 * sin(phi2) * sin(phi2));
                    const double q0 = albers_q(e, sin(phi0)), q1 = albers_q(e, sin(phi1)), q2 = albers_q(e, sin(phi2));
                    const double nn = (m1 * m1 - m2 * m2) / (q2 - q1);
                    const double Cc = m1 * m1 + nn * q1;
                    const double rho0 = a * sqrt(Cc - nn * q0) / nn;
                    const double q = albers_q(e, sin(v0));
                    const double rho = a * sqrt(Cc - nn * q) / nn;
                    const double theta = nn * (v1 - lam0);
                    v0 = x0 + rho * sin(theta);
                    v1 = y0 + rho0 - rho * cos(theta);
                    break;
                }
                case PB200_PROJ_SET_Z: v2 = z_in; break;
                case PB200_PROJ_WEBMERC_FWD: {
                    const double R = 6378137.0;
                    const double lat = v0 * (PB_PI / 180.0), lon = v1 * (PB_PI / 180.0);
                    v0 = R * lon;
                    v1 = R * log(tan(PB_PI / 4.0 + lat / 2.0));
                    break;
                }
                case PB200_PROJ_TMERC_FWD: {  // Snyder 8-9..8-15
                    const double a = p[0], f = 1.0 / p[1], e2 = f * (2.0 - f), ep2 = e2 / (1.0 - e2);
                    const double lat0 = p[2], lon0 = p[3], k0 = p[4], x0 = p[5], y0 = p[6];
                    const double phi = v0, lam = v1;
                    const double sp = sin(phi), cp = cos(phi), tp = tan(phi);
                    const double N = a / sqrt(1.0 - e2 * sp * sp);
                    const double T = tp * tp, Cq = ep2 * cp * cp, A = (lam - lon0) * cp;
                    const double Mv = tmerc_M(a, e2, phi), M0 = tmerc_M(a, e2, lat0);
                    const double A2 = A * A, A3 = A2 * A, A4 = A3 * A, A5 = A4 * A, A6 = A5 * A;
                    v0 = x0 + k0 * N * (A + (1.0 - T + Cq) * A3 / 6.0 + (5.0 - 18.0 * T + T * T + 72.0 * Cq - 58.0 * ep2) * A5 / 120.0);
                    v1 = y0 + k0 * (Mv - M0 + N * tp * (A2 / 2.0 + (5.0 - T + 9.0 * Cq + 4.0 * Cq * Cq) * A4 / 24.0 +
                                                        (61.0 - 58.0 * T + T * T + 600.0 * Cq - 330.0 * ep2) * A6 / 720.0));
                    break;
                }
                default: break;
            }
        }
        double* out = reinterpret_cast<double*>(dst + i * dstride);
        out[0] = v0; out[1] = v1; out[2] = v2;
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_proj_pipeline_for_crs(const char* source_crs, const char* target_crs, pb200_proj_op* ops, uint32_t cap) {
    if (!source_crs || !target_crs || !ops) return set_error(PB200_ERR_INVALID, "null argument");
    const std::string s(source_crs), t(target_crs);
    if (s == "EPSG:4326" && t == "EPSG:3309") {
        if (cap < 5) return set_error(PB200_ERR_INVALID, "ops capacity too small");
        memset(ops, 0, 5 * sizeof(pb200_proj_op));
        ops[0].kind = PB200_PROJ_GEODETIC_TO_ECEF; ops[0].p[0] = 6378137.0; ops[0].p[1] = 298.257223563;  // WGS84, axis order lat, lon
        ops[1].kind = PB200_PROJ_AFFINE;  // inverse of the NAD27 -> WGS84 shift (-8, +159, +175)
        ops[1].p[0] = 1.0; ops[1].p[4] = 1.0; ops[1].p[8] = 1.0; ops[1].p[9] = 8.0; ops[1].p[10] = -159.0; ops[1].p[11] = -175.0;
        ops[2].kind = PB200_PROJ_ECEF_TO_GEODETIC; ops[2].p[0] = 6378206.4; ops[2].p[1] = 294.978698213898;  // Clarke 1866
        ops[3].kind = PB200_PROJ_ALBERS_FWD; ops[3].p[0] = 6378206.4; ops[3].p[1] = 294.978698213898;
        ops[3].p[2] = 34.0 * PB_PI / 180.0; ops[3].p[3] = 40.5 * PB_PI / 180.0; ops[3].p[4] = 0.0;
        ops[3].p[5] = -120.0 * PB_PI / 180.0; ops[3].p[6] = 0.0; ops[3].p[7] = -4000000.0;
        ops[4].kind = PB200_PROJ_SET_Z;
        return 5;
    }
    if (s == "EPSG:4326" && t == "EPSG:3857") {
        if (cap < 1) return set_error(PB200_ERR_INVALID, "ops capacity too small");
        memset(ops, 0, sizeof(pb200_proj_op));
        ops[0].kind = PB200_PROJ_WEBMERC_FWD;
        return 1;
    }
    return set_error(PB200_ERR_UNSUPPORTED, "no built-in pipeline for %s -> %s (general PROJ strings are not supported on the GPU path)",
                     source_crs, target_crs);
}

int pb200_reproject(pb200_ctx* ctx, const pb200_buffer_desc* src, const pb200_buffer_desc* dst_or_null,
                    const pb200_proj_op* ops, uint32_t n_ops) {
    if (!ctx || !ops) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(src, "source point cloud"));
    const pb200_buffer_desc* dst = dst_or_null ? dst_or_null : src;
    if (dst_or_null) PB_TRY(validate_desc(dst, "target point cloud"));
    if (src->len != dst->len) return set_error(PB200_ERR_RANGE, "The point clouds don't have the same size!");  // reprojection.rs:212-214
    if (n_ops > (uint32_t)MAX_PROJ_OPS) return set_error(PB200_ERR_INVALID, "too many pipeline operations");
    PB_DEVICE(ctx);
    const int si = pb200_layout_index_of(src->layout, "Position3D", PB200_VEC3F64);
    const int di = pb200_layout_index_of(dst->layout, "Position3D", PB200_VEC3F64);
    if (si < 0 || di < 0) return set_error(PB200_ERR_ATTR_NOT_FOUND, "buffer has no Vec3f64 Position3D attribute");
    const uint64_t n = src->len;
    if (n == 0) return PB200_OK;
    auto view = [](const pb200_buffer_desc* b, int idx, uint64_t* stride, uint64_t* off) -> uint8_t* {
        const pb200_attr& a = b->layout->attrs[(size_t)idx];
        if (b->kind == PB200_INTERLEAVED) { *stride = b->layout->size; *off = a.offset; return (uint8_t*)b->aos; }
        *stride = a.size; *off = 0;
        return (uint8_t*)b->columns[idx];
    };
    uint64_t ss, so, ds, doff;
    uint8_t* sp = view(src, si, &ss, &so);
    uint8_t* dp = view(dst, di, &ds, &doff);
    ProjPipeline pl;
    pl.n = n_ops;
    for (uint32_t k = 0; k < n_ops; ++k) pl.ops[k] = ops[k];
    // host buffers: stage the position streams
    void *d_s = nullptr, *d_d = nullptr;
    auto cleanup = [&]() { if (d_s) cudaFree(d_s); if (d_d) cudaFree(d_d); };
    const uint8_t* ksrc = sp + so;
    uint8_t* kdst = dp + doff;
    uint64_t kss = ss, kds = ds;
    if (src->memspace == PB200_HOST) {
        // gather positions into a packed device array (24 B/point)
        std::vector<double> packed((size_t)n * 3);
        for (uint64_t i = 0; i < n; ++i) memcpy(&packed[(size_t)i * 3], sp + so + i * ss, 24);
        PB_CUDA(cudaMalloc(&d_s, (size_t)n * 24));
        cudaError_t e = cudaMemcpyAsync(d_s, packed.data(), (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cleanup(); return cuda_error(e, "H2D"); }
        ksrc = (const uint8_t*)d_s;
        kss = 24;
    }
    if (dst->memspace == PB200_HOST) {
        cudaError_t e = cudaMalloc(&d_d, (size_t)n * 24);
        if (e != cudaSuccess) { cleanup(); return cuda_error(e, "cudaMalloc"); }
        kdst = (uint8_t*)d_d;
        kds = 24;
    }
    if (((uintptr_t)ksrc & 7) || (kss & 7) || ((uintptr_t)kdst & 7) || (kds & 7)) {
        cleanup();
        return set_error(PB200_ERR_UNSUPPORTED, "POSITION_3D must be 8-byte aligned in device memory");
    }
    unsigned long long want = (n + 255) / 256, cap = (unsigned long long)ctx->sm_count * 16;
    reproject_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(ksrc, kss, kdst, kds, n, pl);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cleanup(); return cuda_error(e, "reproject_kernel"); }
    if (dst->memspace == PB200_HOST) {
        std::vector<double> packed((size_t)n * 3);
        e = cudaMemcpyAsync(packed.data(), d_d, (size_t)n * 24, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cleanup(); return cuda_error(e, "D2H"); }
        for (uint64_t i = 0; i < n; ++i) memcpy(dp + doff + i * ds, &packed[(size_t)i * 3], 24);
    } else if (src->memspace == PB200_HOST) {
        cudaStreamSynchronize(ctx->stream);
    }
    cleanup();
    return PB200_OK;
}

}  // extern "C"
