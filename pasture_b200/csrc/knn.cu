// knn.cu -- k-nearest-neighbour / radius queries over a device-built linear BVH, and the quirk-faithful
// normal estimation of pasture-algorithms/src/normal_estimation.rs on top of it.
//
// The reference builds a kd-tree with the third-party crate kd-tree 0.3.0 (normal_estimation.rs:103) and asks it for
// `nearests(point, k)` per point (:108).  Tree shape is not part of the contract; the RESULT is: the k points with the
// smallest squared distance ((dx*dx + dy*dy) + dz*dz in f64), the query itself included, ascending.  Here:
//   K10  63-bit Morton codes inside the global AABB (expand_bits_by_3, math/bitmanip.rs:2-10)
//   K8   radix sort of (code, index)
//   K11  Karras-style hierarchy from the sorted codes (ties broken by position in the sorted array) + bottom-up
//        refit of conservative f32 boxes with per-node arrival counters
//   K12  one thread per query in Morton order, stack traversal nearest-child-first, exact f64 distances, a sorted
//        k-list ordered by (d2, original index); a subtree is skipped only if its box distance is > the current worst
//   K13  centroid -> covariance (neighbour order) -> closed-form cubic -> cross products, as written in the reference
//        (the eigenvalue shift at :446-449 is a no-op there)
#include <cub/device/device_radix_sort.cuh>

#include <cfloat>
#include <cmath>

#include "internal.h"

namespace pb200 {

constexpr uint32_t LEAF_FLAG = 0x80000000u;
constexpr int MAX_K = 64;
constexpr int STACK_DEPTH = 128;  // >= 63 code bits + 32 tie-break bits + slack

struct DBuf {
    void* p = nullptr;
    ~DBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

struct Lbvh {
    uint32_t n = 0;
    DBuf codes, codes2, idx, idx2, tmp, spos, left, right, parent, leaf_parent, boxes, counters;
    const double* sorted_pos() const { return (const double*)spos.p; }
};

__device__ __forceinline__ unsigned long long expand3(unsigned long long val) {  // math/bitmanip.rs:2-10
    val &= 0x1FFFFFull;
    val = (val | (val << 32)) & 0x00FF00000000FFFFull;
    val = (val | (val << 16)) & 0x00FF0000FF0000FFull;
    val = (val | (val << 8)) & 0xF00F00F00F00F00Full;
    val = (val | (val << 4)) & 0x30C30C30C30C30C3ull;
    val = (val | (val << 2)) & 0x1249249249249249ull;
    return val;
}

__global__ void __launch_bounds__(256) lbvh_codes_kernel(const uint8_t* __restrict__ base, unsigned long long stride, uint32_t n,
                                                         double bx, double by, double bz, double sx, double sy, double sz,
                                                         unsigned long long* __restrict__ codes, uint32_t* __restrict__ idx) {
    auto quant = [](double p, double b, double s) {
        double t = (p - b) * s;
        if (!(t > 0.0)) return 0ull;
        return t >= 2097151.0 ? 2097151ull : (unsigned long long)__double2ull_rz(t);
    };
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double* p = reinterpret_cast<const double*>(base + (unsigned long long)i * stride);
        codes[i] = (expand3(quant(p[0], bx, sx)) << 2) | (expand3(quant(p[1], by, sy)) << 1) | expand3(quant(p[2], bz, sz));
        idx[i] = i;
    }
}

__global__ void __launch_bounds__(256) gather_positions_kernel(const uint8_t* __restrict__ base, unsigned long long stride,
                                                               const uint32_t* __restrict__ idx, uint32_t n, double* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double* p = reinterpret_cast<const double*>(base + (unsigned long long)idx[i] * stride);
        out[3 * (size_t)i] = p[0];
        out[3 * (size_t)i + 1] = p[1];
        out[3 * (size_t)i + 2] = p[2];
    }
}

__device__ __forceinline__ int delta(const unsigned long long* __restrict__ codes, uint32_t n, long long i, long long j) {
    if (j < 0 || j >= (long long)n) return -1;
    const unsigned long long a = codes[i], b = codes[j];
    if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);  // duplicates: fall back to the position in the sorted array
    return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(256) lbvh_hierarchy_kernel(const unsigned long long* __restrict__ codes, uint32_t n,
                                                             uint32_t* __restrict__ left, uint32_t* __restrict__ right,
                                                             uint32_t* __restrict__ parent, uint32_t* __restrict__ leaf_parent) {
    for (uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 + 1 < n; t0 += gridDim.x * blockDim.x) {
        const long long i = t0;
        const int d = (delta(codes, n, i, i + 1) - delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
        const int dmin = delta(codes, n, i, i - d);
        long long lmax = 2;
        while (delta(codes, n, i, i + lmax * d) > dmin) lmax *= 2;
        long long l = 0;
        for (long long t = lmax / 2; t >= 1; t /= 2)
            if (delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
        const long long j = i + l * d;
        const int dnode = delta(codes, n, i, j);
        long long s = 0, t = l;
        do {
            t = (t + 1) >> 1;
            if (delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
        } while (t > 1);
        const long long gamma = i + s * d + (d < 0 ? d : 0);
        const long long lo = i < j ? i : j, hi = i < j ? j : i;
        const uint32_t lc = (lo == gamma) ? ((uint32_t)gamma | LEAF_FLAG) : (uint32_t)gamma;
        const uint32_t rc = (hi == gamma + 1) ? ((uint32_t)(gamma + 1) | LEAF_FLAG) : (uint32_t)(gamma + 1);
        left[i] = lc;
        right[i] = rc;
        if (lc & LEAF_FLAG) leaf_parent[lc & ~LEAF_FLAG] = (uint32_t)i; else parent[lc] = (uint32_t)i;
        if (rc & LEAF_FLAG) leaf_parent[rc & ~LEAF_FLAG] = (uint32_t)i; else parent[rc] = (uint32_t)i;
        if (i == 0) parent[0] = 0xFFFFFFFFu;
    }
}

struct Box { float lo[3], hi[3]; };

__device__ __forceinline__ Box child_box(uint32_t c, const double* __restrict__ spos, const Box* __restrict__ boxes) {
    if (c & LEAF_FLAG) {
        const double* p = spos + 3 * (size_t)(c & ~LEAF_FLAG);
        Box b;
        for (int a = 0; a < 3; ++a) { b.lo[a] = __double2float_rd(p[a]); b.hi[a] = __double2float_ru(p[a]); }
        return b;
    }
    // written by another SM during this launch: read through L2 (a neighbouring box may sit in a stale L1 line)
    const float* f = reinterpret_cast<const float*>(&boxes[c]);
    Box b;
    for (int a = 0; a < 3; ++a) { b.lo[a] = __ldcg(f + a); b.hi[a] = __ldcg(f + 3 + a); }
    return b;
}

__global__ void __launch_bounds__(256) lbvh_refit_kernel(uint32_t n, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                                                         const uint32_t* __restrict__ parent, const uint32_t* __restrict__ leaf_parent,
                                                         const double* __restrict__ spos, Box* boxes, uint32_t* counters) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t node = leaf_parent[i];
        while (node != 0xFFFFFFFFu) {
            if (atomicAdd(&counters[node], 1u) == 0u) break;  // the second arrival owns the node
            __threadfence();
            const Box a = child_box(left[node], spos, boxes), b = child_box(right[node], spos, boxes);
            Box u;
            for (int c = 0; c < 3; ++c) { u.lo[c] = fminf(a.lo[c], b.lo[c]); u.hi[c] = fmaxf(a.hi[c], b.hi[c]); }
            // volatile-style publication: write, then fence before the parent's counter is touched
            boxes[node] = u;
            __threadfence();
            node = parent[node];
        }
    }
}

__device__ __forceinline__ double box_dist2(const Box& b, double qx, double qy, double qz) {
    const double dx = fmax(fmax((double)b.lo[0] - qx, 0.0), qx - (double)b.hi[0]);
    const double dy = fmax(fmax((double)b.lo[1] - qy, 0.0), qy - (double)b.hi[1]);
    const double dz = fmax(fmax((double)b.lo[2] - qz, 0.0), qz - (double)b.hi[2]);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// ---- normal estimation on an explicit neighbourhood (normal_estimation.rs:198-476) ----------------------------
__device__ void solve_quadratic(double c2, double c1, double ev[3]) {  // :308-325
    ev[0] = 0.0;
    double delta = __dsub_rn(__dmul_rn(c2, c2), __dmul_rn(4.0, c1));
    if (delta < 0.0) delta = 0.0;
    const double sd = sqrt(delta);
    ev[2] = __dmul_rn(0.5, __dadd_rn(c2, sd));
    ev[1] = __dmul_rn(0.5, __dsub_rn(c2, sd));
}

__device__ void solve_polynomial(const double C[9], double ev[3]) {  // :328-392, every product / sum rounded separately
#define M(r, c) C[(r) * 3 + (c)]
    const double t1 = __dmul_rn(__dmul_rn(M(0, 0), M(1, 1)), M(2, 2));
    const double t2 = __dmul_rn(__dmul_rn(__dmul_rn(2.0, M(0, 1)), M(0, 2)), M(1, 2));
    const double t3 = __dmul_rn(__dmul_rn(M(0, 0), M(1, 2)), M(1, 2));
    const double t4 = __dmul_rn(__dmul_rn(M(1, 1), M(0, 2)), M(0, 2));
    const double t5 = __dmul_rn(__dmul_rn(M(2, 2), M(0, 1)), M(0, 1));
    const double c0 = __dsub_rn(__dsub_rn(__dsub_rn(__dadd_rn(t1, t2), t3), t4), t5);
    const double u1 = __dmul_rn(M(0, 0), M(1, 1)), u2 = __dmul_rn(M(0, 1), M(0, 1)), u3 = __dmul_rn(M(0, 0), M(2, 2)),
                 u4 = __dmul_rn(M(0, 2), M(0, 2)), u5 = __dmul_rn(M(1, 1), M(2, 2)), u6 = __dmul_rn(M(1, 2), M(1, 2));
    const double c1 = __dsub_rn(__dadd_rn(__dsub_rn(__dadd_rn(__dsub_rn(u1, u2), u3), u4), u5), u6);
    const double c2 = __dadd_rn(__dadd_rn(M(0, 0), M(1, 1)), M(2, 2));
#undef M
    if (fabs(c0) < 2.220446049250313e-16) { solve_quadratic(c2, c1, ev); return; }
    const double one_third = 1.0 / 3.0;
    const double sqrt_3 = sqrt(3.0);
    const double c2_third = __dmul_rn(c2, one_third);
    double alpha_third = __dmul_rn(__dsub_rn(c1, __dmul_rn(c2, c2_third)), one_third);
    if (alpha_third > 0.0) alpha_third = 0.0;
    const double half_beta =
        __dmul_rn(0.5, __dadd_rn(c0, __dmul_rn(c2_third, __dsub_rn(__dmul_rn(__dmul_rn(2.0, c2_third), c2_third), c1))));
    double q = __dadd_rn(__dmul_rn(half_beta, half_beta), __dmul_rn(__dmul_rn(alpha_third, alpha_third), alpha_third));
    if (q > 0.0) q = 0.0;
    const double rho = sqrt(-alpha_third);
    const double theta = __dmul_rn(atan2(sqrt(-q), half_beta), one_third);
    const double ct = cos(theta), st = sin(theta);
    ev[0] = __dadd_rn(c2_third, __dmul_rn(__dmul_rn(2.0, rho), ct));
    ev[1] = __dsub_rn(c2_third, __dmul_rn(rho, __dadd_rn(ct, __dmul_rn(sqrt_3, st))));
    ev[2] = __dsub_rn(c2_third, __dmul_rn(rho, __dsub_rn(ct, __dmul_rn(sqrt_3, st))));
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2 - i; ++j)
            if (ev[j + 1] < ev[j]) { double t = ev[j]; ev[j] = ev[j + 1]; ev[j + 1] = t; }
    if (ev[0] <= 0.0) solve_quadratic(c2, c1, ev);
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* r) {
    r[0] = __dsub_rn(__dmul_rn(a[1], b[2]), __dmul_rn(a[2], b[1]));
    r[1] = __dsub_rn(__dmul_rn(a[2], b[0]), __dmul_rn(a[0], b[2]));
    r[2] = __dsub_rn(__dmul_rn(a[0], b[1]), __dmul_rn(a[1], b[0]));
}
__device__ __forceinline__ double norm3(const double* a) {
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a[0], a[0]), __dmul_rn(a[1], a[1])), __dmul_rn(a[2], a[2])));
}

// neighbours: cnt original indices in kNN order; positions fetched through (base, stride)
__device__ void estimate_normal(const uint8_t* __restrict__ base, unsigned long long stride, const uint32_t* nb, uint32_t cnt,
                                double normal[3], double* curvature) {
    // compute_centroid :198-237 (dense case)
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (uint32_t j = 0; j < cnt; ++j) {
        const double* p = reinterpret_cast<const double*>(base + (unsigned long long)nb[j] * stride);
        t0 = __dadd_rn(t0, p[0]); t1 = __dadd_rn(t1, p[1]); t2 = __dadd_rn(t2, p[2]);
    }
    const double c0 = t0 / (double)cnt, c1 = t1 / (double)cnt, c2 = t2 / (double)cnt;
    // compute_covariance_matrix :240-305
    double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t j = 0; j < cnt; ++j) {
        const double* p = reinterpret_cast<const double*>(base + (unsigned long long)nb[j] * stride);
        const double d0 = __dsub_rn(p[0], c0), d1 = __dsub_rn(p[1], c1), d2 = __dsub_rn(p[2], c2);
        C[4] = __dadd_rn(C[4], __dmul_rn(d1, d1));
        C[5] = __dadd_rn(C[5], __dmul_rn(d1, d2));
        C[8] = __dadd_rn(C[8], __dmul_rn(d2, d2));
        C[0] = __dadd_rn(C[0], __dmul_rn(d0, d0));
        C[1] = __dadd_rn(C[1], __dmul_rn(d1, d0));
        C[2] = __dadd_rn(C[2], __dmul_rn(d2, d0));
    }
    C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
    // eigen_3x3 :429-453
    double scale = 0.0;
    for (int i = 0; i < 9; ++i) { const double a = fabs(C[i]); if (a > scale) scale = a; }
    double S[9];
    for (int i = 0; i < 9; ++i) S[i] = C[i] / scale;
    double ev[3];
    solve_polynomial(C, ev);
    const double eigen_value = __dmul_rn(ev[0], scale);
    double rows[3][3];
    cross3(&S[0], &S[3], rows[0]);
    cross3(&S[0], &S[6], rows[1]);
    cross3(&S[3], &S[6], rows[2]);
    int best = 0;
    for (int r = 1; r < 3; ++r)
        if (norm3(rows[r]) > norm3(rows[best])) best = r;
    normal[0] = rows[best][0]; normal[1] = rows[best][1]; normal[2] = rows[best][2];
    const double eigen_sum = __dadd_rn(__dadd_rn(C[0], C[4]), C[8]);
    *curvature = eigen_sum != 0.0 ? fabs(eigen_value / eigen_sum) : 0.0;
}

struct QueryArgs {
    uint32_t n, k;
    const double* spos;        // positions in Morton order
    const uint32_t* sidx;      // original index of sorted position i
    const uint32_t *left, *right;
    const Box* boxes;
    double radius2;            // < 0: pure kNN
    uint32_t* idx_out;         // n*k (original order), nullable
    double* d2_out;            // n*k, nullable
    uint32_t* counts_out;      // n, nullable (radius search)
    const uint8_t* pos_base;   // original-order positions (normals)
    unsigned long long pos_stride;
    double* normals_out;       // n*3, nullable
    double* curvature_out;     // n, nullable
};

// MODE 0: kNN lists, 1: radius search, 2: normals
template <int MODE>
__global__ void __launch_bounds__(128) lbvh_query_kernel(QueryArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const double qx = a.spos[3 * (size_t)i], qy = a.spos[3 * (size_t)i + 1], qz = a.spos[3 * (size_t)i + 2];
    const uint32_t self = a.sidx[i];
    const uint32_t k = a.k;
    double bd[MAX_K];
    uint32_t bi[MAX_K];
    uint32_t cnt = 0;
    const double limit = MODE == 1 ? a.radius2 : DBL_MAX;
    auto worst = [&]() { return cnt == k ? bd[k - 1] : limit; };
    auto offer = [&](double d2, uint32_t id) {
        if (!(d2 <= limit)) return;  // also rejects NaN distances
        if (cnt == k && !(d2 < bd[k - 1] || (d2 == bd[k - 1] && id < bi[k - 1]))) return;
        uint32_t pos = cnt < k ? cnt : k - 1;
        while (pos > 0 && (d2 < bd[pos - 1] || (d2 == bd[pos - 1] && id < bi[pos - 1]))) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
        bd[pos] = d2;
        bi[pos] = id;
        if (cnt < k) ++cnt;
    };
    if (a.n == 1) {
        offer(0.0, self);
    } else {
        uint32_t stack[STACK_DEPTH];
        int sp = 0;
        uint32_t node = 0;  // root
        while (true) {
            const uint32_t cl = a.left[node], cr = a.right[node];
            double dl, dr;
            if (cl & LEAF_FLAG) {
                const double* p = a.spos + 3 * (size_t)(cl & ~LEAF_FLAG);
                const double dx = p[0] - qx, dy = p[1] - qy, dz = p[2] - qz;
                dl = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                offer(dl, a.sidx[cl & ~LEAF_FLAG]);
                dl = DBL_MAX;  // consumed
            } else dl = box_dist2(a.boxes[cl], qx, qy, qz);
            if (cr & LEAF_FLAG) {
                const double* p = a.spos + 3 * (size_t)(cr & ~LEAF_FLAG);
                const double dx = p[0] - qx, dy = p[1] - qy, dz = p[2] - qz;
                dr = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                offer(dr, a.sidx[cr & ~LEAF_FLAG]);
                dr = DBL_MAX;
            } else dr = box_dist2(a.boxes[cr], qx, qy, qz);
            const bool vl = !(cl & LEAF_FLAG) && dl <= worst(), vr = !(cr & LEAF_FLAG) && dr <= worst();
            if (vl && vr) {
                const bool left_first = dl <= dr;
                if (sp < STACK_DEPTH) stack[sp++] = left_first ? cr : cl;
                node = left_first ? cl : cr;
            } else if (vl) node = cl;
            else if (vr) node = cr;
            else {
                // pop, re-checking the bound: the list may have tightened since the push
                bool found = false;
                while (sp > 0) {
                    const uint32_t c = stack[--sp];
                    if (box_dist2(a.boxes[c], qx, qy, qz) <= worst()) { node = c; found = true; break; }
                }
                if (!found) break;
            }
        }
    }
    if (MODE == 2) {
        double nrm[3], curv;
        estimate_normal(a.pos_base, a.pos_stride, bi, cnt, nrm, &curv);
        a.normals_out[3 * (size_t)self] = nrm[0];
        a.normals_out[3 * (size_t)self + 1] = nrm[1];
        a.normals_out[3 * (size_t)self + 2] = nrm[2];
        a.curvature_out[self] = curv;
        return;
    }
    for (uint32_t j = 0; j < k; ++j) {
        if (a.idx_out) a.idx_out[(size_t)self * k + j] = j < cnt ? bi[j] : 0xFFFFFFFFu;
        if (a.d2_out) a.d2_out[(size_t)self * k + j] = j < cnt ? bd[j] : INFINITY;
    }
    if (a.counts_out) a.counts_out[self] = cnt;
}

static unsigned grid_for(uint64_t n, int sm, unsigned block = 256) {
    uint64_t want = (n + block - 1) / block, cap = (uint64_t)sm * 32;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

// positions of `buf` as a device (base, stride) view; host buffers are staged
static int device_positions(pb200_ctx* ctx, const pb200_buffer_desc* buf, DBuf* staged, const uint8_t** base, uint64_t* stride) {
    const int pi = pb200_layout_index_of(buf->layout, "Position3D", PB200_VEC3F64);
    if (pi < 0) return set_error(PB200_ERR_ATTR_NOT_FOUND, "buffer has no Vec3f64 Position3D attribute (view_attribute::<Vector3<f64>> would panic)");
    const pb200_attr& at = buf->layout->attrs[(size_t)pi];
    const uint8_t* p;
    if (buf->kind == PB200_INTERLEAVED) { *stride = buf->layout->size; p = (const uint8_t*)buf->aos; }
    else { *stride = at.size; p = (const uint8_t*)buf->columns[pi]; }
    if (buf->memspace == PB200_HOST) {
        const size_t bytes = (size_t)(buf->len * (*stride));
        PB_CUDA(staged->alloc(bytes));
        PB_CUDA(cudaMemcpyAsync(staged->p, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
        p = (const uint8_t*)staged->p;
    }
    *base = p + (buf->kind == PB200_INTERLEAVED ? at.offset : 0);
    if (((uintptr_t)*base & 7) || (*stride & 7)) return set_error(PB200_ERR_UNSUPPORTED, "POSITION_3D must be 8-byte aligned in memory");
    return PB200_OK;
}

static int build_lbvh(pb200_ctx* ctx, const uint8_t* base, uint64_t stride, uint32_t n, Lbvh* t) {
    cudaStream_t st = ctx->stream;
    t->n = n;
    // global AABB on the device-resident view (an SoA-like descriptor over the strided positions)
    pb200_layout l;
    pb200_attr a{};
    strcpy(a.name, "Position3D");
    a.dtype = PB200_VEC3F64;
    a.size = 24;
    a.offset = 0;
    l.attrs.push_back(a);
    l.size = stride;
    l.align = 1;
    pb200_buffer_desc d;
    d.layout = &l;
    d.kind = PB200_INTERLEAVED;
    d.memspace = PB200_DEVICE;
    d.len = n;
    d.aos = (void*)base;
    d.columns = nullptr;
    double bmin[3], bmax[3];
    int some = 0;
    PB_TRY(pb200_calculate_bounds(ctx, &d, bmin, bmax, &some));
    double s[3];
    for (int c = 0; c < 3; ++c) { const double e = bmax[c] - bmin[c]; s[c] = e > 0.0 ? 2097152.0 / e : 0.0; }
    PB_CUDA(t->codes.alloc((size_t)n * 8)); PB_CUDA(t->codes2.alloc((size_t)n * 8));
    PB_CUDA(t->idx.alloc((size_t)n * 4)); PB_CUDA(t->idx2.alloc((size_t)n * 4));
    PB_CUDA(t->spos.alloc((size_t)n * 24));
    lbvh_codes_kernel<<<grid_for(n, ctx->sm_count), 256, 0, st>>>(base, stride, n, bmin[0], bmin[1], bmin[2], s[0], s[1], s[2],
                                                                  (unsigned long long*)t->codes.p, (uint32_t*)t->idx.p);
    g_launches++;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned long long*)t->codes.p, (unsigned long long*)t->codes2.p,
                                    (const uint32_t*)t->idx.p, (uint32_t*)t->idx2.p, (int)n, 0, 63, st);
    PB_CUDA(t->tmp.alloc(tmp_bytes));
    PB_CUDA(cub::DeviceRadixSort::SortPairs(t->tmp.p, tmp_bytes, (const unsigned long long*)t->codes.p, (unsigned long long*)t->codes2.p,
                                            (const uint32_t*)t->idx.p, (uint32_t*)t->idx2.p, (int)n, 0, 63, st));
    g_launches += 9;
    gather_positions_kernel<<<grid_for(n, ctx->sm_count), 256, 0, st>>>(base, stride, (const uint32_t*)t->idx2.p, n, (double*)t->spos.p);
    g_launches++;
    if (n >= 2) {
        PB_CUDA(t->left.alloc((size_t)(n - 1) * 4)); PB_CUDA(t->right.alloc((size_t)(n - 1) * 4));
        PB_CUDA(t->parent.alloc((size_t)(n - 1) * 4)); PB_CUDA(t->leaf_parent.alloc((size_t)n * 4));
        PB_CUDA(t->boxes.alloc((size_t)(n - 1) * sizeof(Box))); PB_CUDA(t->counters.alloc((size_t)(n - 1) * 4));
        PB_CUDA(cudaMemsetAsync(t->counters.p, 0, (size_t)(n - 1) * 4, st));
        lbvh_hierarchy_kernel<<<grid_for(n - 1, ctx->sm_count), 256, 0, st>>>((const unsigned long long*)t->codes2.p, n, (uint32_t*)t->left.p,
                                                                              (uint32_t*)t->right.p, (uint32_t*)t->parent.p, (uint32_t*)t->leaf_parent.p);
        g_launches++;
        lbvh_refit_kernel<<<grid_for(n, ctx->sm_count), 256, 0, st>>>(n, (const uint32_t*)t->left.p, (const uint32_t*)t->right.p,
                                                                      (const uint32_t*)t->parent.p, (const uint32_t*)t->leaf_parent.p,
                                                                      (const double*)t->spos.p, (Box*)t->boxes.p, (uint32_t*)t->counters.p);
        g_launches++;
    }
    PB_CUDA(cudaGetLastError());
    return PB200_OK;
}

// run a query kernel; outputs may live in host memory (staged through device temporaries)
static int run_query(pb200_ctx* ctx, const pb200_buffer_desc* buf, int mode, uint32_t k, double radius, uint32_t* idx_out,
                     double* d2_out, uint32_t* counts_out, double* normals_out, double* curvature_out) {
    PB_TRY(validate_desc(buf, "buffer"));
    PB_TRY(ensure_device(ctx));
    if (buf->len > 0x7FFFFFFFull) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^31-1 points per call");
    if (k == 0 || k > MAX_K) return set_error(PB200_ERR_UNSUPPORTED, "k must be in 1..%d", MAX_K);
    const uint32_t n = (uint32_t)buf->len;
    if (n == 0) return PB200_OK;
    DBuf staged;
    const uint8_t* base = nullptr;
    uint64_t stride = 0;
    PB_TRY(device_positions(ctx, buf, &staged, &base, &stride));
    Lbvh tree;
    PB_TRY(build_lbvh(ctx, base, stride, n, &tree));
    const bool host = buf->memspace == PB200_HOST;
    DBuf d_idx, d_d2, d_cnt, d_nrm, d_curv;
    QueryArgs a{};
    a.n = n; a.k = k;
    a.spos = tree.sorted_pos();
    a.sidx = (const uint32_t*)tree.idx2.p;
    a.left = (const uint32_t*)tree.left.p; a.right = (const uint32_t*)tree.right.p;
    a.boxes = (const Box*)tree.boxes.p;
    a.radius2 = mode == 1 ? radius * radius : -1.0;
    a.pos_base = base; a.pos_stride = stride;
    auto out_ptr = [&](void* user, DBuf& tmp, size_t bytes, void** dev) -> int {
        *dev = nullptr;
        if (!user) return PB200_OK;
        if (!host) { *dev = user; return PB200_OK; }
        PB_CUDA(tmp.alloc(bytes));
        *dev = tmp.p;
        return PB200_OK;
    };
    void* p = nullptr;
    PB_TRY(out_ptr(idx_out, d_idx, (size_t)n * k * 4, &p)); a.idx_out = (uint32_t*)p;
    PB_TRY(out_ptr(d2_out, d_d2, (size_t)n * k * 8, &p)); a.d2_out = (double*)p;
    PB_TRY(out_ptr(counts_out, d_cnt, (size_t)n * 4, &p)); a.counts_out = (uint32_t*)p;
    PB_TRY(out_ptr(normals_out, d_nrm, (size_t)n * 24, &p)); a.normals_out = (double*)p;
    PB_TRY(out_ptr(curvature_out, d_curv, (size_t)n * 8, &p)); a.curvature_out = (double*)p;
    const unsigned blocks = (n + 127) / 128;
    if (mode == 0) lbvh_query_kernel<0><<<blocks, 128, 0, ctx->stream>>>(a);
    else if (mode == 1) lbvh_query_kernel<1><<<blocks, 128, 0, ctx->stream>>>(a);
    else lbvh_query_kernel<2><<<blocks, 128, 0, ctx->stream>>>(a);
    g_launches++;
    PB_CUDA(cudaGetLastError());
    if (host) {
        if (idx_out) PB_CUDA(cudaMemcpyAsync(idx_out, a.idx_out, (size_t)n * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (d2_out) PB_CUDA(cudaMemcpyAsync(d2_out, a.d2_out, (size_t)n * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (counts_out) PB_CUDA(cudaMemcpyAsync(counts_out, a.counts_out, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (normals_out) PB_CUDA(cudaMemcpyAsync(normals_out, a.normals_out, (size_t)n * 24, cudaMemcpyDeviceToHost, ctx->stream));
        if (curvature_out) PB_CUDA(cudaMemcpyAsync(curvature_out, a.curvature_out, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    // the tree and staging buffers die with this call: wait for the kernels that use them
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_knn(pb200_ctx* ctx, const pb200_buffer_desc* buf, uint32_t k, uint32_t* idx_out, double* d2_out) {
    if (!ctx || !idx_out) return set_error(PB200_ERR_INVALID, "null argument");
    return run_query(ctx, buf, 0, k, 0.0, idx_out, d2_out, nullptr, nullptr, nullptr);
}

int pb200_radius_search(pb200_ctx* ctx, const pb200_buffer_desc* buf, double radius, uint32_t max_neighbors,
                        uint32_t* idx_out, uint32_t* counts_out) {
    if (!ctx || !idx_out || !counts_out) return set_error(PB200_ERR_INVALID, "null argument");
    if (!(radius >= 0.0)) return set_error(PB200_ERR_INVALID, "radius must be >= 0");
    return run_query(ctx, buf, 1, max_neighbors, radius, idx_out, nullptr, counts_out, nullptr, nullptr);
}

int pb200_compute_normals(pb200_ctx* ctx, const pb200_buffer_desc* buf, uint32_t k, double* normals_out,
                          double* curvature_out) {
    if (!ctx || !buf || !normals_out || !curvature_out) return set_error(PB200_ERR_INVALID, "null argument");
    if (buf->len < 3)  // normal_estimation.rs:86-88
        return set_error(PB200_ERR_TOO_FEW_POINTS, "The point cloud is too small. Please use a point cloud that has 3 or more points!");
    if (k < 3)  // :89-91
        return set_error(PB200_ERR_INVALID, "The k nearest neigbors attribute is too small!");
    return run_query(ctx, buf, 2, k, 0.0, nullptr, nullptr, nullptr, normals_out, curvature_out);
}

}  // extern "C"
