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
                case PB200_PROJ_TMERC_FWD: {
                    // Transverse Mercator, Krueger series in the third flattening n (the "JHS" formulas of IOGP Guidance
                    // Note 7-2, EPSG method 9807; sub-millimetre within +-4 deg of the central meridian, and what PROJ's
                    // default tmerc/utm evaluates).  p = compiled on the host: e, B*k0, h1..h4, lon0, FE, FN - k0*M0.
                    const double e = p[0], Bk0 = p[1], lam0 = p[6];
                    const double Q = asinh(tan(v0)) - e * atanh(e * sin(v0));
                    const double beta = atan(sinh(Q));
                    const double eta0 = atanh(cos(beta) * sin(v1 - lam0));
                    const double xi0 = asin(sin(beta) * cosh(eta0));
                    double xi = xi0, eta = eta0;
#pragma unroll
                    for (int j = 1; j <= 4; ++j) {
                        xi += p[1 + j] * sin(2.0 * j * xi0) * cosh(2.0 * j * eta0);
                        eta += p[1 + j] * cos(2.0 * j * xi0) * sinh(2.0 * j * eta0);
                    }
                    v0 = p[7] + Bk0 * eta;
                    v1 = p[8] + Bk0 * xi;
                    break;
                }
                case PB200_PROJ_TMERC_INV: {  // GN7-2 reverse formulas; p: e, B*k0, h1'..h4', lon0, FE, FN - k0*M0
                    const double e = p[0], Bk0 = p[1], lam0 = p[6];
                    const double eta = (v0 - p[7]) / Bk0, xi = (v1 - p[8]) / Bk0;
                    double xi0 = xi, eta0 = eta;
#pragma unroll
                    for (int j = 1; j <= 4; ++j) {
                        xi0 -= p[1 + j] * sin(2.0 * j * xi) * cosh(2.0 * j * eta);
                        eta0 -= p[1 + j] * cos(2.0 * j * xi) * sinh(2.0 * j * eta);
                    }
                    const double beta = asin(sin(xi0) / cosh(eta0));
                    const double Qp = asinh(tan(beta));
                    double Q = Qp;
                    for (int it = 0; it < 8; ++it) Q = Qp + e * atanh(e * tanh(Q));  // converges by ~e^2 per step
                    v0 = atan(sinh(Q));
                    v1 = lam0 + asin(tanh(eta0) / cos(beta));
                    break;
                }
                case PB200_PROJ_DEG2RAD_LATLON: v0 *= PB_PI / 180.0; v1 *= PB_PI / 180.0; break;
                case PB200_PROJ_RAD2DEG_LATLON: v0 *= 180.0 / PB_PI; v1 *= 180.0 / PB_PI; break;
                case PB200_PROJ_WEBMERC_INV: {  // EPSG method 1024 reverse: (E, N) -> (lat_deg, lon_deg)
                    const double R = 6378137.0;
                    const double lon = v0 / R, lat = PB_PI / 2.0 - 2.0 * atan(exp(-v1 / R));
                    v0 = lat * (180.0 / PB_PI);
                    v1 = lon * (180.0 / PB_PI);
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

namespace {

// Transverse Mercator series constants (IOGP Guidance Note 7-2, 3.5.3.1): public op parameters (a, 1/f, lat0, lon0, k0,
// FE, FN; radians) -> what the kernel evaluates per point (e, B*k0, four series coefficients, lon0, FE, FN - k0*M0)
void compile_tmerc(const pb200_proj_op& in, bool inverse, pb200_proj_op* out) {
    const double a = in.p[0], f = 1.0 / in.p[1], lat0 = in.p[2], lon0 = in.p[3], k0 = in.p[4], fe = in.p[5], fn = in.p[6];
    const double n = f / (2.0 - f), n2 = n * n, n3 = n2 * n, n4 = n2 * n2, e = sqrt(f * (2.0 - f));
    const double B = a / (1.0 + n) * (1.0 + n2 / 4.0 + n4 / 64.0);
    const double h[4] = {n / 2.0 - 2.0 / 3.0 * n2 + 5.0 / 16.0 * n3 + 41.0 / 180.0 * n4, 13.0 / 48.0 * n2 - 3.0 / 5.0 * n3 + 557.0 / 1440.0 * n4,
                         61.0 / 240.0 * n3 - 103.0 / 140.0 * n4, 49561.0 / 161280.0 * n4};
    const double hi[4] = {n / 2.0 - 2.0 / 3.0 * n2 + 37.0 / 96.0 * n3 - 1.0 / 360.0 * n4, 1.0 / 48.0 * n2 + 1.0 / 15.0 * n3 - 437.0 / 1440.0 * n4,
                          17.0 / 480.0 * n3 - 37.0 / 840.0 * n4, 4397.0 / 161280.0 * n4};
    double M0 = 0.0;
    if (lat0 != 0.0) {
        const double Q0 = asinh(tan(lat0)) - e * atanh(e * sin(lat0));
        const double xi00 = atan(sinh(Q0));
        double xi = xi00;
        for (int j = 1; j <= 4; ++j) xi += h[j - 1] * sin(2.0 * j * xi00);
        M0 = B * xi;
    }
    memset(out, 0, sizeof(*out));
    out->kind = in.kind;
    out->p[0] = e;
    out->p[1] = B * k0;
    for (int j = 0; j < 4; ++j) out->p[2 + j] = inverse ? hi[j] : h[j];
    out->p[6] = lon0;
    out->p[7] = fe;
    out->p[8] = fn - k0 * M0;
}

void tmerc_op(pb200_proj_op* op, bool inverse, double a, double inv_f, double lat0_deg, double lon0_deg, double k0, double fe, double fn) {
    memset(op, 0, sizeof(*op));
    op->kind = inverse ? PB200_PROJ_TMERC_INV : PB200_PROJ_TMERC_FWD;
    op->p[0] = a; op->p[1] = inv_f; op->p[2] = lat0_deg * PB_PI / 180.0; op->p[3] = lon0_deg * PB_PI / 180.0;
    op->p[4] = k0; op->p[5] = fe; op->p[6] = fn;
}

// EPSG codes of projected CRSs the pipeline builder knows: UTM on WGS 84 (326zz north, 327zz south) and on ETRS89
// (258zz, zones 28-38; GRS 1980).  ETRS89 and WGS 84 coincide at the accuracy PROJ's default operation assumes.
struct UtmCrs { bool ok; double a, inv_f; int zone; bool south; };
UtmCrs parse_utm(const std::string& crs) {
    UtmCrs u{false, 0, 0, 0, false};
    if (crs.size() != 10 || crs.compare(0, 5, "EPSG:") != 0) return u;
    const int code = atoi(crs.c_str() + 5);
    if (code >= 32601 && code <= 32660) u = {true, 6378137.0, 298.257223563, code - 32600, false};
    else if (code >= 32701 && code <= 32760) u = {true, 6378137.0, 298.257223563, code - 32700, true};
    else if (code >= 25828 && code <= 25838) u = {true, 6378137.0, 298.257222101, code - 25800, false};
    return u;
}
void utm_op(pb200_proj_op* op, bool inverse, const UtmCrs& u) {
    tmerc_op(op, inverse, u.a, u.inv_f, 0.0, -183.0 + 6.0 * u.zone, 0.9996, 500000.0, u.south ? 10000000.0 : 0.0);
}

}  // namespace

extern "C" {

int pb200_proj_op_tmerc(double a, double inv_f, double lat0_deg, double lon0_deg, double k0, double false_easting, double false_northing,
                        int inverse, pb200_proj_op* out) {
    if (!out) return set_error(PB200_ERR_INVALID, "null argument");
    if (!(a > 0.0) || !(inv_f > 1.0) || !(k0 > 0.0)) return set_error(PB200_ERR_INVALID, "bad ellipsoid / scale factor");
    tmerc_op(out, inverse != 0, a, inv_f, lat0_deg, lon0_deg, k0, false_easting, false_northing);
    return PB200_OK;
}

int pb200_proj_op_helmert(double tx, double ty, double tz, double rx_arcsec, double ry_arcsec, double rz_arcsec, double ds_ppm,
                          int coordinate_frame, pb200_proj_op* out) {
    if (!out) return set_error(PB200_ERR_INVALID, "null argument");
    const double as = PB_PI / (180.0 * 3600.0), sgn = coordinate_frame ? -1.0 : 1.0;  // EPSG 1032 = 1033 with the rotations negated
    const double rx = sgn * rx_arcsec * as, ry = sgn * ry_arcsec * as, rz = sgn * rz_arcsec * as, m = 1.0 + ds_ppm * 1e-6;
    memset(out, 0, sizeof(*out));
    out->kind = PB200_PROJ_AFFINE;  // GN7-2 4.3.3: XT = M * (1 -rz +ry; +rz 1 -rx; -ry +rx 1) * XS + t
    out->p[0] = m;        out->p[1] = -m * rz;  out->p[2] = m * ry;
    out->p[3] = m * rz;   out->p[4] = m;        out->p[5] = -m * rx;
    out->p[6] = -m * ry;  out->p[7] = m * rx;   out->p[8] = m;
    out->p[9] = tx; out->p[10] = ty; out->p[11] = tz;
    return PB200_OK;
}

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
    if (s == "EPSG:3857" && t == "EPSG:4326") {
        if (cap < 1) return set_error(PB200_ERR_INVALID, "ops capacity too small");
        memset(ops, 0, sizeof(pb200_proj_op));
        ops[0].kind = PB200_PROJ_WEBMERC_INV;
        return 1;
    }
    {   // UTM zones on WGS 84 / ETRS89 <-> geographic WGS 84 (lat, lon in degrees: EPSG:4326 axis order) and zone <-> zone
        const UtmCrs us = parse_utm(s), ut = parse_utm(t);
        if (s == "EPSG:4326" && ut.ok) {
            if (cap < 2) return set_error(PB200_ERR_INVALID, "ops capacity too small");
            memset(ops, 0, 2 * sizeof(pb200_proj_op));
            ops[0].kind = PB200_PROJ_DEG2RAD_LATLON;
            utm_op(&ops[1], false, ut);
            return 2;
        }
        if (us.ok && t == "EPSG:4326") {
            if (cap < 2) return set_error(PB200_ERR_INVALID, "ops capacity too small");
            memset(ops, 0, 2 * sizeof(pb200_proj_op));
            utm_op(&ops[0], true, us);
            ops[1].kind = PB200_PROJ_RAD2DEG_LATLON;
            return 2;
        }
        if (us.ok && ut.ok) {
            if (cap < 2) return set_error(PB200_ERR_INVALID, "ops capacity too small");
            utm_op(&ops[0], true, us);
            utm_op(&ops[1], false, ut);
            return 2;
        }
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
    for (uint32_t k = 0; k < n_ops; ++k) {
        pl.ops[k] = ops[k];
        if (ops[k].kind == PB200_PROJ_TMERC_FWD || ops[k].kind == PB200_PROJ_TMERC_INV) {
            if (!(ops[k].p[0] > 0.0) || !(ops[k].p[1] > 1.0) || !(ops[k].p[4] > 0.0))
                return set_error(PB200_ERR_INVALID, "transverse Mercator op %u: bad ellipsoid / scale factor", k);
            compile_tmerc(ops[k], ops[k].kind == PB200_PROJ_TMERC_INV, &pl.ops[k]);
        }
    }
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
