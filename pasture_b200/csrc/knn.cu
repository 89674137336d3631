// knn.cu -- k-nearest-neighbour / radius queries over a device-built linear BVH, and the quirk-faithful
// normal estimation of pasture-algorithms/src/normal_estimation.rs on top of it.
//
// The reference builds a kd-tree with the third-party crate kd-tree 0.3.0 (normal_estimation.rs:103) and asks it for
// `nearests(point, k)` per point (:108).  Tree shape is not part of the contract; the RESULT is: the k points with the
// smallest squared distance ((dx*dx + dy*dy) + dz*dz in f64), the query itself included, ascending.  Here:
//   K10  63-bit Morton codes inside the global AABB (expand_bits_by_3, math/bitmanip.rs:2-10)
//   K8   radix sort of (code, index) (radix_sort.cu); positions gathered into Morton order (padded with +inf to whole buckets)
//   K11  BUCKETS of 8 consecutive sorted points are the leaves; a Karras-style radix hierarchy is built over the first
//        code of every bucket (ties broken by bucket number).  One 64-byte record per internal node holds BOTH child
//        boxes (conservative f32, rounded outwards) and the split, so a visit is four 16-byte loads and decides on both
//        children at once; boxes are refitted bottom-up with per-node arrival counters.
//   K12  one thread per query, queries in Morton order (a warp walks neighbouring buckets, so node and bucket loads hit
//        L1).  The k-list lives in REGISTERS (template KMAX, worst element in slot 0, fully unrolled insertion) ordered
//        by (d2, original index); it is primed from the query's own and adjacent buckets before the top-down,
//        nearest-child-first traversal, so the pruning bound is tight from the first node on.  A subtree is skipped
//        only if its box distance (same f64 association as the point distance, hence monotone) is > the current worst.
//   K13  centroid -> covariance (neighbour order) -> closed-form cubic -> cross products, as written in the reference
//        (the eigenvalue shift at :446-449 is a no-op there); fused into the query kernel.
#include <cfloat>
#include <cmath>
#include <type_traits>

#include "internal.h"

namespace pb200 {

constexpr int BUCKET = 8;                 // points per leaf
constexpr uint32_t LEFT_LEAF = 0x80000000u, RIGHT_LEAF = 0x40000000u, SPLIT_MASK = 0x3FFFFFFFu;
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int MAX_K = 64;
constexpr int STACK_DEPTH = 128;          // >= 63 code bits + 28 tie-break bits + slack

struct alignas(16) Node {                 // 64 B: what one traversal step needs
    float l[6];                           // left child box: lo xyz, hi xyz
    float r[6];                           // right child box
    uint32_t split;                       // gamma | LEFT_LEAF | RIGHT_LEAF; children are gamma and gamma + 1
    uint32_t first, last;                 // bucket range covered by this node
    uint32_t pad;
};
static_assert(sizeof(Node) == 64, "node record is one 64-byte line");

struct Lbvh {
    uint32_t n = 0, nb = 0;
    double origin[3] = {0.0, 0.0, 0.0};
    DevTmp codes, codes2, idx, idx2, tmp, spos, nodes, parent, leaf_parent, counters;
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

// positions in Morton order; the tail of the last bucket is +inf (such a point is never accepted: d2 = inf)
__global__ void __launch_bounds__(256) gather_positions_kernel(const uint8_t* __restrict__ base, unsigned long long stride,
                                                               const uint32_t* __restrict__ idx, uint32_t n, uint32_t n_padded,
                                                               double* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_padded; i += gridDim.x * blockDim.x) {
        double x = INFINITY, y = INFINITY, z = INFINITY;
        if (i < n) {
            // random 24-byte gathers: 64-byte L2 fills (LDG.E.LTC64B) instead of the default granularity
            const uint8_t* p = base + (unsigned long long)idx[i] * stride;
            asm volatile("ld.global.nc.L2::64B.f64 %0, [%1];" : "=d"(x) : "l"(p));
            asm volatile("ld.global.nc.L2::64B.f64 %0, [%1];" : "=d"(y) : "l"(p + 8));
            asm volatile("ld.global.nc.L2::64B.f64 %0, [%1];" : "=d"(z) : "l"(p + 16));
        }
        out[3 * (size_t)i] = x;
        out[3 * (size_t)i + 1] = y;
        out[3 * (size_t)i + 2] = z;
    }
}

// common-prefix length of the keys of buckets i and j (key = code of the bucket's first point)
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ codes, uint32_t nb, long long i, long long j) {
    if (j < 0 || j >= (long long)nb) return -1;
    const unsigned long long a = codes[(size_t)i * BUCKET], b = codes[(size_t)j * BUCKET];
    if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);  // duplicates: fall back to the bucket number
    return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(256) lbvh_hierarchy_kernel(const unsigned long long* __restrict__ codes, uint32_t nb,
                                                             Node* __restrict__ nodes, uint32_t* __restrict__ parent,
                                                             uint32_t* __restrict__ leaf_parent) {
    for (uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 + 1 < nb; t0 += gridDim.x * blockDim.x) {
        const long long i = t0;
        const int d = (delta(codes, nb, i, i + 1) - delta(codes, nb, i, i - 1)) >= 0 ? 1 : -1;
        const int dmin = delta(codes, nb, i, i - d);
        long long lmax = 2;
        while (delta(codes, nb, i, i + lmax * d) > dmin) lmax *= 2;
        long long l = 0;
        for (long long t = lmax / 2; t >= 1; t /= 2)
            if (delta(codes, nb, i, i + (l + t) * d) > dmin) l += t;
        const long long j = i + l * d;
        const int dnode = delta(codes, nb, i, j);
        long long s = 0, t = l;
        do {
            t = (t + 1) >> 1;
            if (delta(codes, nb, i, i + (s + t) * d) > dnode) s += t;
        } while (t > 1);
        const long long gamma = i + s * d + (d < 0 ? d : 0);
        const long long lo = i < j ? i : j, hi = i < j ? j : i;
        const bool ll = lo == gamma, rl = hi == gamma + 1;
        nodes[i].split = (uint32_t)gamma | (ll ? LEFT_LEAF : 0u) | (rl ? RIGHT_LEAF : 0u);
        nodes[i].first = (uint32_t)lo;
        nodes[i].last = (uint32_t)hi;
        if (ll) leaf_parent[gamma] = (uint32_t)i; else parent[gamma] = (uint32_t)i;
        if (rl) leaf_parent[gamma + 1] = (uint32_t)i; else parent[gamma + 1] = (uint32_t)i;
        if (i == 0) parent[0] = NONE;
    }
}

// one thread per bucket: box of its points, then climb; the second arrival at a node owns it
// f32 interval arithmetic for the box tests.  Boxes are stored RELATIVE to the AABB minimum (georeferenced clouds keep
// sub-millimetre box resolution) and widened: round outwards, then one more f32 ulp, which swallows the f64 rounding of
// the translation itself.  Together with round-down arithmetic and the safety factor in box_lower_bound this makes the
// f32 value a strict lower bound of the f64 point distance the k-list compares against.
__device__ __forceinline__ float f32_below(double v) { return nextafterf(__double2float_rd(v), -INFINITY); }
__device__ __forceinline__ float f32_above(double v) { return nextafterf(__double2float_ru(v), INFINITY); }

__global__ void __launch_bounds__(256) lbvh_refit_kernel(uint32_t n, uint32_t nb, Node* nodes, const uint32_t* __restrict__ parent,
                                                         const uint32_t* __restrict__ leaf_parent,
                                                         const double* __restrict__ spos, double ox, double oy, double oz,
                                                         uint32_t* counters) {
    const double origin[3] = {ox, oy, oz};
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        const uint32_t p0 = b * BUCKET, p1 = (p0 + BUCKET < n) ? p0 + BUCKET : n;
        for (uint32_t p = p0; p < p1; ++p)
            for (int a = 0; a < 3; ++a) {
                const double v = spos[3 * (size_t)p + a];
                if (v != v) continue;  // NaN coordinates never become neighbours; keep them out of the boxes
                lo[a] = fminf(lo[a], f32_below(v - origin[a]));
                hi[a] = fmaxf(hi[a], f32_above(v - origin[a]));
            }
        uint32_t child = b, node = leaf_parent[b];
        while (node != NONE) {
            Node* nd = nodes + node;
            const bool is_left = child == (nd->split & SPLIT_MASK);  // the left child (leaf or node) is number gamma
            float* mine = is_left ? nd->l : nd->r;
            for (int a = 0; a < 3; ++a) { __stcg(mine + a, lo[a]); __stcg(mine + 3 + a, hi[a]); }
            __threadfence();  // publish the box before the arrival is counted
            if (atomicAdd(&counters[node], 1u) == 0u) break;
            __threadfence();
            const float* other = is_left ? nd->r : nd->l;  // written by another SM during this launch: read through L2
            for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], __ldcg(other + a)); hi[a] = fmaxf(hi[a], __ldcg(other + 3 + a)); }
            child = node;
            node = parent[node];
        }
    }
}

// lower bound of the squared distance between the query interval [ql, qh] (per axis) and a box, all in round-down f32
__device__ __forceinline__ float box_lower_bound(float lx, float ly, float lz, float hx, float hy, float hz, float qlx, float qly,
                                                 float qlz, float qhx, float qhy, float qhz) {
    const float dx = fmaxf(fmaxf(__fsub_rd(lx, qhx), __fsub_rd(qlx, hx)), 0.0f);
    const float dy = fmaxf(fmaxf(__fsub_rd(ly, qhy), __fsub_rd(qly, hy)), 0.0f);
    const float dz = fmaxf(fmaxf(__fsub_rd(lz, qhz), __fsub_rd(qlz, hz)), 0.0f);
    const float d2 = __fadd_rd(__fadd_rd(__fmul_rd(dx, dx), __fmul_rd(dy, dy)), __fmul_rd(dz, dz));
    return __fmul_rd(d2, 0.99999976f);  // 1 - 2^-22: strictly below the f64-rounded point distances
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
__device__ void estimate_normal(const uint8_t* __restrict__ base, unsigned long long stride, const uint32_t* nb, uint32_t nb_stride,
                                uint32_t cnt, double normal[3], double* curvature) {
    // compute_centroid :198-237 (dense case)
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (uint32_t j = 0; j < cnt; ++j) {
        const double* p = reinterpret_cast<const double*>(base + (unsigned long long)nb[j * nb_stride] * stride);
        t0 = __dadd_rn(t0, p[0]); t1 = __dadd_rn(t1, p[1]); t2 = __dadd_rn(t2, p[2]);
    }
    const double c0 = t0 / (double)cnt, c1 = t1 / (double)cnt, c2 = t2 / (double)cnt;
    // compute_covariance_matrix :240-305
    double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t j = 0; j < cnt; ++j) {
        const double* p = reinterpret_cast<const double*>(base + (unsigned long long)nb[j * nb_stride] * stride);
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
    uint32_t n, nb, k, init_radius;
    uint32_t nq, out_base;     // queries answered by this launch / original index of the first one (outputs are relative to it)
    const uint32_t* qpos;      // null: every point is a query (thread t <-> sorted position t); else the sorted positions of
                               // the queries, ascending (= Morton order), one per thread
    const double* spos;        // positions in Morton order, padded to nb * BUCKET points
    const uint32_t* sidx;      // original index of sorted position i
    const Node* nodes;
    double radius2;            // < 0: pure kNN
    double origin[3];          // boxes and query intervals are relative to this point (the AABB minimum)
    uint32_t* idx_out;         // n*k (original order), nullable
    double* d2_out;            // n*k, nullable
    uint32_t* counts_out;      // n, nullable (radius search)
    double* normals_out;       // n*3, nullable
    double* curvature_out;     // n, nullable
};

// The k best candidates, ordered by (d2, original index), worst first: slot 0 is the pruning bound, slots >= k are
// unused.  Empty slots hold (+inf, NONE), which every real candidate beats.  All indices are compile-time constants
// after unrolling, so the list stays in registers.
template <int KMAX>
struct KList {
    double d[KMAX];
    uint32_t j[KMAX];  // SORTED position of the neighbour (its original index is sidx[j], needed for ties only)
    __device__ __forceinline__ void init() {
#pragma unroll KMAX <= 32 ? KMAX : 1
        for (int s = 0; s < KMAX; ++s) { d[s] = INFINITY; j[s] = NONE; }
    }
    __device__ __forceinline__ double worst() const { return d[0]; }
    __device__ __forceinline__ void offer(double d2, uint32_t jj, uint32_t k, double limit, const uint32_t* __restrict__ sidx) {
        if (!(d2 <= limit)) return;  // also rejects NaN and the +inf padding
        if (d2 > d[0]) return;
        if (d2 == d[0] && !(sidx[jj] < sidx[j[0]])) return;
        if constexpr (KMAX > 32) {  // too long for registers: plain insertion in local memory
            uint32_t s = 0;
            while (s + 1 < k && (d[s + 1] > d2 || (d[s + 1] == d2 && sidx[j[s + 1]] > sidx[jj]))) {
                d[s] = d[s + 1];
                j[s] = j[s + 1];
                ++s;
            }
            d[s] = d2;
            j[s] = jj;
            return;
        }
        // Branch-free insertion.  c[s] = "slot s is in use and worse than the candidate"; it holds for a prefix of the
        // slots (the list is sorted, worst first; c[0] holds by the tests above).  The candidate lands in the last
        // slot of that prefix, everything before it moves one slot towards 0, the old slot 0 drops out:
        //     new[s] = c[s+1] ? old[s+1] : (c[s] ? candidate : old[s])
        // Exact distance ties need the original indices (two loads each): handled by the same formula with a tie-aware
        // predicate, on a path that is only taken when some slot really ties.
        bool tie = false;
#pragma unroll
        for (int s = 1; s < KMAX; ++s) tie = tie || d[s] == d2;
        if (!tie) {
            bool cs = true;
#pragma unroll
            for (int s = 0; s < KMAX; ++s) {
                const bool cn = (s + 1 < KMAX) && (uint32_t)(s + 1) < k && d[s + 1 < KMAX ? s + 1 : s] > d2;
                const double nd = d[s + 1 < KMAX ? s + 1 : s];
                const uint32_t nj = j[s + 1 < KMAX ? s + 1 : s];
                d[s] = cn ? nd : (cs ? d2 : d[s]);
                j[s] = cn ? nj : (cs ? jj : j[s]);
                cs = cn;
            }
        } else {
            const uint32_t my = sidx[jj];
            bool cs = true;
#pragma unroll
            for (int s = 0; s < KMAX; ++s) {
                const double nd = d[s + 1 < KMAX ? s + 1 : s];
                const uint32_t nj = j[s + 1 < KMAX ? s + 1 : s];
                bool cn = (s + 1 < KMAX) && (uint32_t)(s + 1) < k;
                if (cn) cn = nd > d2 || (nd == d2 && sidx[nj] > my);
                d[s] = cn ? nd : (cs ? d2 : d[s]);
                j[s] = cn ? nj : (cs ? jj : j[s]);
                cs = cn;
            }
        }
    }
};

// The same k-list as a binary MAX-heap in shared memory (slot s of thread t at s * 128 + t: conflict-free whatever slots
// the lanes touch).  Dynamic indexing costs nothing there, so one instantiation serves every k; replacing the root and
// sifting down touches log2(k) levels instead of all k slots, and the ~50 registers of a 16-entry register list are free
// for occupancy.  Order: (d2, original index), the root is the worst kept candidate = the pruning bound.
struct HeapList {
    double* d;      // this thread's column of the [k][128] arrays
    uint32_t* j;
    uint32_t cnt;
    double bound;   // d[0] once k candidates are kept, +inf before
    static constexpr uint32_t STRIDE = 128;
    __device__ __forceinline__ double& D(uint32_t s) { return d[s * STRIDE]; }
    __device__ __forceinline__ uint32_t& J(uint32_t s) { return j[s * STRIDE]; }
    __device__ __forceinline__ void init(uint8_t* smem, uint32_t k) {
        d = reinterpret_cast<double*>(smem) + threadIdx.x;
        j = reinterpret_cast<uint32_t*>(smem + (size_t)k * STRIDE * 8) + threadIdx.x;
        cnt = 0;
        bound = INFINITY;
    }
    __device__ __forceinline__ double worst() const { return bound; }
    __device__ __forceinline__ bool greater(double da, uint32_t ja, double db, uint32_t jb, const uint32_t* __restrict__ sidx) const {
        return da > db || (da == db && sidx[ja] > sidx[jb]);
    }
    // put (dv, jv) into the hole at slot i of a heap of `size` entries and restore the heap below it
    __device__ __forceinline__ void sift_down(uint32_t i, double dv, uint32_t jv, uint32_t size, const uint32_t* __restrict__ sidx) {
        while (true) {
            uint32_t c = 2 * i + 1;
            if (c >= size) break;
            double dc = D(c);
            uint32_t jc = J(c);
            if (c + 1 < size) {
                const double d2c = D(c + 1);
                const uint32_t j2c = J(c + 1);
                if (greater(d2c, j2c, dc, jc, sidx)) { ++c; dc = d2c; jc = j2c; }
            }
            if (!greater(dc, jc, dv, jv, sidx)) break;
            D(i) = dc; J(i) = jc;
            i = c;
        }
        D(i) = dv; J(i) = jv;
    }
    __device__ __forceinline__ void heapify(uint32_t size, const uint32_t* __restrict__ sidx) {
        for (int i = (int)(size / 2) - 1; i >= 0; --i) sift_down((uint32_t)i, D((uint32_t)i), J((uint32_t)i), size, sidx);
    }
    __device__ __forceinline__ void offer(double d2, uint32_t jj, uint32_t k, double limit, const uint32_t* __restrict__ sidx) {
        if (!(d2 <= limit)) return;  // also rejects NaN and the +inf padding
        if (cnt < k) {               // filling: append, build the heap when the k-th candidate arrives
            D(cnt) = d2; J(cnt) = jj;
            if (++cnt == k) { heapify(k, sidx); bound = D(0); }
            return;
        }
        if (d2 > bound) return;
        if (d2 == bound && !(sidx[jj] < sidx[J(0)])) return;
        sift_down(0, d2, jj, k, sidx);  // the root (worst) drops out
        bound = D(0);
    }
    // heap -> ascending (d2, index) order in slots [0, cnt)
    __device__ __forceinline__ void finish(uint32_t k, const uint32_t* __restrict__ sidx) {
        if (cnt < k) heapify(cnt, sidx);
        for (uint32_t m = cnt; m > 1; --m) {
            const double dv = D(m - 1);
            const uint32_t jv = J(m - 1);
            D(m - 1) = D(0); J(m - 1) = J(0);
            sift_down(0, dv, jv, m - 1, sidx);
        }
    }
};

// MODE 0: kNN lists, 1: radius search, 2: normals, 3: kNN traversal statistics (diagnostics)
//
// PACKET traversal: the 32 queries of a warp are consecutive in Morton order (four buckets), so their search regions
// overlap almost completely.  The warp walks the tree ONCE: node and bucket addresses are warp-uniform (every load is a
// broadcast), each lane tests the two child boxes against its OWN query and bound, and a child is entered if ANY lane
// needs it (ballot).  Control flow never diverges in the traversal; only the list insertions do.  The result per lane is
// exactly what its private traversal would give: a lane ignores buckets whose box is beyond its own bound because no
// point in them passes its candidate test.
// Control flow is also arranged so that the (fully unrolled, KMAX-long) list insertion is instantiated exactly once:
// buckets to scan are queued as a range [pend_lo, pend_hi) -- the priming range first, then the one or two leaf
// children of the visited node, which are consecutive buckets by construction (gamma, gamma + 1).
// KMAX = 0 selects the shared-memory heap (HeapList, any k), KMAX > 0 the register list of that width.
template <int KMAX, int MODE>
__global__ void __launch_bounds__(128) lbvh_query_kernel(QueryArgs a) {
    constexpr bool HEAP = KMAX == 0;
    constexpr int KREG = HEAP ? 1 : KMAX;
    extern __shared__ __align__(16) uint8_t knn_smem[];
    __shared__ uint32_t s_stack[4][STACK_DEPTH];  // one warp-uniform stack per warp
    const uint32_t t_q = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool active = t_q < a.nq;
    const uint32_t i = active ? (a.qpos ? a.qpos[t_q] : t_q) : 0u;  // sorted position of this thread's query
    // lanes beyond the cloud keep voting but never need anything: NaN fails every comparison
    const double qnan = __longlong_as_double(0x7FF8000000000000ll);
    const double qx = active ? a.spos[3 * (size_t)i] : qnan, qy = active ? a.spos[3 * (size_t)i + 1] : qnan,
                 qz = active ? a.spos[3 * (size_t)i + 2] : qnan;
    const float qlx = f32_below(qx - a.origin[0]), qly = f32_below(qy - a.origin[1]), qlz = f32_below(qz - a.origin[2]);
    const float qhx = f32_above(qx - a.origin[0]), qhy = f32_above(qy - a.origin[1]), qhz = f32_above(qz - a.origin[2]);
    const uint32_t k = a.k;
    const double limit = MODE == 1 ? a.radius2 : DBL_MAX;
    const uint32_t* __restrict__ sidx = a.sidx;
    typename std::conditional<HEAP, HeapList, KList<KREG>>::type list;
    if constexpr (HEAP) list.init(knn_smem, k);
    else list.init();
    // prime the lists from the buckets of the warp's queries and their neighbours in Morton order.  With every point a
    // query these are four consecutive buckets; a query subset (qpos) is still in Morton order but spread out, so the
    // primed window is capped around the warp's middle query and the traversal finds the rest.
    const uint32_t act = __ballot_sync(0xffffffffu, active);
    const uint32_t last_lane = act ? 31u - (uint32_t)__clz((int)act) : 0u;
    const uint32_t qb0 = __shfl_sync(0xffffffffu, i, 0) / BUCKET, qb1 = __shfl_sync(0xffffffffu, i, last_lane) / BUCKET;
    const uint32_t qbm = __shfl_sync(0xffffffffu, i, last_lane / 2) / BUCKET;
    constexpr uint32_t PRIME_HALF = 6;
    uint32_t ib0 = qb0 > a.init_radius ? qb0 - a.init_radius : 0u;
    uint32_t ib1 = (qb1 + a.init_radius < a.nb - 1) ? qb1 + a.init_radius : a.nb - 1;
    if (ib1 - ib0 + 1 > 2 * PRIME_HALF) {
        ib0 = qbm > PRIME_HALF ? qbm - PRIME_HALF : 0u;
        ib1 = (qbm + PRIME_HALF - 1 < a.nb - 1) ? qbm + PRIME_HALF - 1 : a.nb - 1;
    }
    uint32_t pend_lo = ib0, pend_hi = ib1 + 1;
    bool more = a.nb > 1 && !(ib0 == 0 && ib1 == a.nb - 1);
    int sp = 0;
    uint32_t node = 0;  // root
    uint32_t st_nodes = 0, st_buckets = 0, st_offers = 0;
    while (true) {
        for (uint32_t b = pend_lo; b < pend_hi; ++b) {
            if (MODE == 3) ++st_buckets;
            const double2* p = reinterpret_cast<const double2*>(a.spos) + (size_t)b * (BUCKET * 3 / 2);
            double c[BUCKET * 3], dd[BUCKET];
#pragma unroll
            for (int t = 0; t < BUCKET * 3 / 2; ++t) { const double2 v = __ldg(p + t); c[2 * t] = v.x; c[2 * t + 1] = v.y; }
            uint32_t cand = 0;
            const double bound = fmin(list.worst(), limit);  // radius search: nothing beyond r^2 is ever needed, full list or not
#pragma unroll
            for (int t = 0; t < BUCKET; ++t) {
                const double dx = c[3 * t] - qx, dy = c[3 * t + 1] - qy, dz = c[3 * t + 2] - qz;
                dd[t] = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                cand |= (dd[t] <= bound) ? (1u << t) : 0u;
            }
            while (cand) {
                const int t = __ffs((int)cand) - 1;
                cand &= cand - 1;
                double d2 = dd[0];
#pragma unroll
                for (int u = 1; u < BUCKET; ++u) d2 = (t == u) ? dd[u] : d2;
                if (MODE == 3) ++st_offers;
                list.offer(d2, b * BUCKET + (uint32_t)t, k, limit, sidx);
            }
            __syncwarp();
        }
        pend_lo = pend_hi = 0;
        if (!more) break;
        if (MODE == 3) ++st_nodes;
        const float4* rec = reinterpret_cast<const float4*>(a.nodes + node);
        const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
        const uint32_t split = __ldg(reinterpret_cast<const uint32_t*>(rec + 3));
        const float dl = box_lower_bound(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, qlx, qly, qlz, qhx, qhy, qhz);
        const float dr = box_lower_bound(r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, qlx, qly, qlz, qhx, qhy, qhz);
        const uint32_t cl = split & SPLIT_MASK, cr = cl + 1;
        const bool ll = (split & LEFT_LEAF) != 0, rl = (split & RIGHT_LEAF) != 0;
        const double bound = fmin(list.worst(), limit);  // radius search: nothing beyond r^2 is ever needed, full list or not
        const bool need_l = active && (double)dl <= bound, need_r = active && (double)dr <= bound;
        const uint32_t bl = __ballot_sync(0xffffffffu, need_l), br = __ballot_sync(0xffffffffu, need_r);
        const uint32_t left_votes = __popc(__ballot_sync(0xffffffffu, need_l && (!need_r || dl <= dr)));
        const uint32_t right_votes = __popc(__ballot_sync(0xffffffffu, need_r && (!need_l || dr < dl)));
        const bool sl = ll && bl != 0 && (cl < ib0 || cl > ib1), sr = rl && br != 0 && (cr < ib0 || cr > ib1);
        if (sl || sr) { pend_lo = sl ? cl : cr; pend_hi = sr ? cr + 1 : cl + 1; }
        const bool vl = !ll && bl != 0, vr = !rl && br != 0;
        if (vl && vr) {
            const bool left_first = left_votes >= right_votes;
            if (sp < STACK_DEPTH) {
                if (lane == 0) s_stack[warp][sp] = left_first ? cr : cl;
                ++sp;
            }
            node = left_first ? cl : cr;
        } else if (vl) node = cl;
        else if (vr) node = cr;
        else if (sp > 0) {
            __syncwarp();  // lane 0's pushes are visible
            node = s_stack[warp][--sp];
            __syncwarp();  // everyone has read the entry before lane 0 may overwrite it
        } else more = false;
    }
    if (!active) return;
    const uint32_t self = sidx[i] - a.out_base;  // output slot: original index relative to the first query
    if (MODE == 3) {  // (nodes visited, buckets scanned, candidates offered) per query
        a.idx_out[(size_t)self * 3] = st_nodes;
        a.idx_out[(size_t)self * 3 + 1] = st_buckets;
        a.idx_out[(size_t)self * 3 + 2] = st_offers;
        return;
    }
    if constexpr (HEAP) {
        list.finish(k, sidx);  // ascending (d2, index) order = the order kd-tree's nearests() returns
        if (MODE == 2) {
            double nrm[3], curv;
            estimate_normal(reinterpret_cast<const uint8_t*>(a.spos), 24ull, list.j, HeapList::STRIDE, list.cnt, nrm, &curv);
            a.normals_out[3 * (size_t)self] = nrm[0];
            a.normals_out[3 * (size_t)self + 1] = nrm[1];
            a.normals_out[3 * (size_t)self + 2] = nrm[2];
            a.curvature_out[self] = curv;
            return;
        }
        for (uint32_t s = 0; s < k; ++s) {
            const bool valid = s < list.cnt;
            if (a.idx_out) a.idx_out[(size_t)self * k + s] = valid ? sidx[list.J(s)] : NONE;
            if (a.d2_out) a.d2_out[(size_t)self * k + s] = valid ? list.D(s) : INFINITY;
        }
        if (a.counts_out) a.counts_out[self] = list.cnt;
    } else {
        if (MODE == 2) {
            uint32_t nb[KREG];  // ascending (d2, index) order = the order kd-tree's nearests() returns
            uint32_t cnt = 0;
#pragma unroll KREG <= 32 ? KREG : 1
            for (int s = KREG - 1; s >= 0; --s)
                if ((uint32_t)s < k && list.j[s] != NONE) nb[cnt++] = list.j[s];
            double nrm[3], curv;
            estimate_normal(reinterpret_cast<const uint8_t*>(a.spos), 24ull, nb, 1u, cnt, nrm, &curv);
            a.normals_out[3 * (size_t)self] = nrm[0];
            a.normals_out[3 * (size_t)self + 1] = nrm[1];
            a.normals_out[3 * (size_t)self + 2] = nrm[2];
            a.curvature_out[self] = curv;
            return;
        }
        uint32_t cnt = 0;
#pragma unroll KREG <= 32 ? KREG : 1
        for (int s = 0; s < KREG; ++s) {
            if ((uint32_t)s < k) {
                const uint32_t o = k - 1u - (uint32_t)s;
                const bool valid = list.j[s] != NONE;
                cnt += valid ? 1u : 0u;
                if (a.idx_out) a.idx_out[(size_t)self * k + o] = valid ? sidx[list.j[s]] : NONE;
                if (a.d2_out) a.d2_out[(size_t)self * k + o] = list.d[s];
            }
        }
        if (a.counts_out) a.counts_out[self] = cnt;
    }
}

// ---- query subsets: the sorted positions of the points whose ORIGINAL index lies in [first, first + nq), ascending -------
// (ordered stream compaction of the sorted index array: per-tile counts -> scan -> emit; a thread owns 8 consecutive slots)
constexpr int QC_THREADS = 256, QC_PER = 8, QC_TILE = QC_THREADS * QC_PER;

__global__ void __launch_bounds__(QC_THREADS) qpos_count_kernel(const uint32_t* __restrict__ sidx, uint32_t n, uint32_t first, uint32_t nq,
                                                                uint32_t* __restrict__ tile_counts) {
    __shared__ uint32_t warp_sum[QC_THREADS / 32];
    const uint32_t b = blockIdx.x * QC_TILE + threadIdx.x * QC_PER;
    uint32_t c = 0;
#pragma unroll
    for (int u = 0; u < QC_PER; ++u)
        if (b + u < n) c += (sidx[b + u] - first) < nq ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < QC_THREADS / 32; ++w) t += warp_sum[w];
        tile_counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(QC_THREADS) qpos_emit_kernel(const uint32_t* __restrict__ sidx, uint32_t n, uint32_t first, uint32_t nq,
                                                               const uint32_t* __restrict__ tile_offsets, uint32_t* __restrict__ qpos) {
    __shared__ uint32_t warp_sum[QC_THREADS / 32];
    const uint32_t b = blockIdx.x * QC_TILE + threadIdx.x * QC_PER, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t flags = 0;
#pragma unroll
    for (int u = 0; u < QC_PER; ++u)
        if (b + u < n && (sidx[b + u] - first) < nq) flags |= 1u << u;
    const uint32_t c = __popc(flags);
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((int)lane >= o) x += y; }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    uint32_t before = tile_offsets[blockIdx.x];
    for (uint32_t w = 0; w < warp; ++w) before += warp_sum[w];
    uint32_t o = before + x - c;
#pragma unroll
    for (int u = 0; u < QC_PER; ++u)
        if ((flags >> u) & 1u) qpos[o++] = b + (uint32_t)u;
}

static unsigned grid_for(uint64_t n, int sm, unsigned block = 256) {
    uint64_t want = (n + block - 1) / block, cap = (uint64_t)sm * 32;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

// positions of `buf` as a device (base, stride) view; host buffers are staged
static int device_positions(pb200_ctx* ctx, const pb200_buffer_desc* buf, DevTmp* staged, const uint8_t** base, uint64_t* stride) {
    const int pi = pb200_layout_index_of(buf->layout, "Position3D", PB200_VEC3F64);
    if (pi < 0) return set_error(PB200_ERR_ATTR_NOT_FOUND, "buffer has no Vec3f64 Position3D attribute (view_attribute::<Vector3<f64>> would panic)");
    const pb200_attr& at = buf->layout->attrs[(size_t)pi];
    const uint8_t* p;
    if (buf->kind == PB200_INTERLEAVED) { *stride = buf->layout->size; p = (const uint8_t*)buf->aos; }
    else { *stride = at.size; p = (const uint8_t*)buf->columns[pi]; }
    if (buf->memspace == PB200_HOST) {
        const size_t bytes = (size_t)(buf->len * (*stride));
        PB_CUDA(staged->alloc(ctx->stream, bytes));
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
    t->nb = (n + BUCKET - 1) / BUCKET;
    const uint32_t nb = t->nb, n_padded = nb * BUCKET;
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
    { PB_PHASE(ctx, "knn.bounds"); PB_TRY(pb200_calculate_bounds(ctx, &d, bmin, bmax, &some)); }
    // Quantisation for the tree's Morton order: ONE scale for all axes (cubic cells), so that neighbours on the curve
    // are neighbours in space even for 2.5-D clouds whose z extent is a fraction of x/y.  (pb200_morton_codes, the
    // public Z-row entry point, quantises per axis inside the AABB.)
    for (int c = 0; c < 3; ++c) t->origin[c] = (some && bmin[c] == bmin[c] && fabs(bmin[c]) <= DBL_MAX) ? bmin[c] : 0.0;
    double s[3];
    double emax = 0.0;
    for (int c = 0; c < 3; ++c) { const double e = bmax[c] - bmin[c]; if (e > emax) emax = e; }
    for (int c = 0; c < 3; ++c) {
        const double e = ctx->knn_per_axis_codes ? bmax[c] - bmin[c] : emax;
        s[c] = e > 0.0 ? 2097152.0 / e : 0.0;
    }
    PB_CUDA(t->codes.alloc(st, (size_t)n * 8)); PB_CUDA(t->codes2.alloc(st, (size_t)n * 8));
    PB_CUDA(t->idx.alloc(st, (size_t)n * 4)); PB_CUDA(t->idx2.alloc(st, (size_t)n * 4));
    PB_CUDA(t->spos.alloc(st, (size_t)n_padded * 24));
    {
        PB_PHASE(ctx, "knn.codes");
        lbvh_codes_kernel<<<grid_for(n, ctx->sm_count), 256, 0, st>>>(base, stride, n, bmin[0], bmin[1], bmin[2], s[0], s[1], s[2],
                                                                      (unsigned long long*)t->codes.p, (uint32_t*)t->idx.p);
        g_launches++;
    }
    {   // K8: own one-sweep radix sort over the 63 code bits; the sorted codes / indices must end up in codes2 / idx2
        bool in_alt = false;
        PB_TRY(radix_sort_u64(ctx, (unsigned long long*)t->codes.p, (unsigned long long*)t->codes2.p, (uint32_t*)t->idx.p,
                              (uint32_t*)t->idx2.p, n, 0, 63, &in_alt));
        if (!in_alt) { std::swap(t->codes.p, t->codes2.p); std::swap(t->idx.p, t->idx2.p); }
    }
    {
        PB_PHASE(ctx, "knn.gather_positions");
        gather_positions_kernel<<<grid_for(n_padded, ctx->sm_count), 256, 0, st>>>(base, stride, (const uint32_t*)t->idx2.p, n, n_padded,
                                                                                   (double*)t->spos.p);
        g_launches++;
    }
    if (nb >= 2) {
        PB_PHASE(ctx, "knn.tree");
        PB_CUDA(t->nodes.alloc(st, (size_t)(nb - 1) * sizeof(Node)));
        PB_CUDA(t->parent.alloc(st, (size_t)(nb - 1) * 4)); PB_CUDA(t->leaf_parent.alloc(st, (size_t)nb * 4));
        PB_CUDA(t->counters.alloc(st, (size_t)(nb - 1) * 4));
        PB_CUDA(cudaMemsetAsync(t->counters.p, 0, (size_t)(nb - 1) * 4, st));
        lbvh_hierarchy_kernel<<<grid_for(nb - 1, ctx->sm_count), 256, 0, st>>>((const unsigned long long*)t->codes2.p, nb, (Node*)t->nodes.p,
                                                                               (uint32_t*)t->parent.p, (uint32_t*)t->leaf_parent.p);
        g_launches++;
        lbvh_refit_kernel<<<grid_for(nb, ctx->sm_count), 256, 0, st>>>(n, nb, (Node*)t->nodes.p, (const uint32_t*)t->parent.p,
                                                                       (const uint32_t*)t->leaf_parent.p, (const double*)t->spos.p,
                                                                       t->origin[0], t->origin[1], t->origin[2], (uint32_t*)t->counters.p);
        g_launches++;
    }
    PB_CUDA(cudaGetLastError());
    return PB200_OK;
}

// the shared-memory heap variant: one instantiation per mode, k * 128 * 12 bytes of dynamic shared memory
static int launch_query_heap(pb200_ctx* ctx, int mode, const QueryArgs& a, cudaStream_t st) {
    const unsigned blocks = (a.nq + 127) / 128;
    const size_t smem = (size_t)a.k * 128 * 12;
    if (!ctx->knn_attr_set) {
        const int max_smem = MAX_K * 128 * 12;
        PB_CUDA(cudaFuncSetAttribute(lbvh_query_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        PB_CUDA(cudaFuncSetAttribute(lbvh_query_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        PB_CUDA(cudaFuncSetAttribute(lbvh_query_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        ctx->knn_attr_set = true;
    }
    if (mode == 0) lbvh_query_kernel<0, 0><<<blocks, 128, smem, st>>>(a);
    else if (mode == 1) lbvh_query_kernel<0, 1><<<blocks, 128, smem, st>>>(a);
    else lbvh_query_kernel<0, 2><<<blocks, 128, smem, st>>>(a);
    return PB200_OK;
}

template <int KMAX>
static void launch_query(int mode, const QueryArgs& a, cudaStream_t st) {
    const unsigned blocks = (a.nq + 127) / 128;
#ifdef PB200_KNN_DIAGNOSTICS
    if (mode == 3) { if constexpr (KMAX == 16) lbvh_query_kernel<16, 3><<<blocks, 128, 0, st>>>(a); return; }
#endif
    if (mode == 0) lbvh_query_kernel<KMAX, 0><<<blocks, 128, 0, st>>>(a);
    else if (mode == 1) lbvh_query_kernel<KMAX, 1><<<blocks, 128, 0, st>>>(a);
    else lbvh_query_kernel<KMAX, 2><<<blocks, 128, 0, st>>>(a);
}

// run a query kernel; outputs may live in host memory (staged through device temporaries)
// queries: the points [first_query, first_query + n_queries) of the buffer (all outputs hold n_queries entries)
static int run_query(pb200_ctx* ctx, const pb200_buffer_desc* buf, int mode, uint32_t k, double radius, uint64_t first_query,
                     uint64_t n_queries, uint32_t* idx_out, double* d2_out, uint32_t* counts_out, double* normals_out,
                     double* curvature_out) {
    PB_TRY(validate_desc(buf, "buffer"));
    PB_DEVICE(ctx);
    if (buf->len > 0x7FFFFFF0ull) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^31-16 points per call");
    if (k == 0 || k > MAX_K) return set_error(PB200_ERR_UNSUPPORTED, "k must be in 1..%d", MAX_K);
    const uint32_t n = (uint32_t)buf->len;
    if (first_query > n || n_queries > n - first_query) return set_error(PB200_ERR_RANGE, "query range %llu + %llu exceeds the %u points of the buffer",
                                                                         (unsigned long long)first_query, (unsigned long long)n_queries, n);
    const uint32_t nq = (uint32_t)n_queries;
    if (n == 0 || nq == 0) return PB200_OK;
    DevTmp staged;
    const uint8_t* base = nullptr;
    uint64_t stride = 0;
    PB_TRY(device_positions(ctx, buf, &staged, &base, &stride));
    Lbvh tree;
    PB_TRY(build_lbvh(ctx, base, stride, n, &tree));
    const bool host = buf->memspace == PB200_HOST;
    DevTmp d_idx, d_d2, d_cnt, d_nrm, d_curv;
    QueryArgs a{};
    a.n = n; a.nb = tree.nb; a.k = k;
    a.nq = nq; a.out_base = (uint32_t)first_query; a.qpos = nullptr;
    DevTmp d_qpos, d_qtiles;
    if (nq != n) {  // a query subset: its sorted positions in Morton order (the replicas-only multi-GPU cut)
        PB_PHASE(ctx, "knn.query_list");
        const uint32_t tiles = (n + QC_TILE - 1) / QC_TILE;
        PB_CUDA(d_qpos.alloc(ctx->stream, (size_t)nq * 4));
        PB_CUDA(d_qtiles.alloc(ctx->stream, ((size_t)tiles + 1) * 4));
        qpos_count_kernel<<<tiles, QC_THREADS, 0, ctx->stream>>>((const uint32_t*)tree.idx2.p, n, a.out_base, nq, (uint32_t*)d_qtiles.p);
        PB_TRY(exclusive_scan_u32(ctx, (uint32_t*)d_qtiles.p, tiles, (uint32_t*)d_qtiles.p + tiles));
        qpos_emit_kernel<<<tiles, QC_THREADS, 0, ctx->stream>>>((const uint32_t*)tree.idx2.p, n, a.out_base, nq, (const uint32_t*)d_qtiles.p,
                                                                (uint32_t*)d_qpos.p);
        g_launches += 2;
        a.qpos = (const uint32_t*)d_qpos.p;
    }
    // the warp's own four buckets +- init_radius buckets prime every lane's list before the traversal
    a.init_radius = ctx->knn_heap ? (k + 15) / 16 : (k + 7) / 8;  // measured per list variant (benchmarks/knn_probe.py)
    if (ctx->knn_init_radius >= 0) a.init_radius = (uint32_t)ctx->knn_init_radius;
    a.spos = tree.sorted_pos();
    a.sidx = (const uint32_t*)tree.idx2.p;
    a.nodes = (const Node*)tree.nodes.p;
    a.radius2 = mode == 1 ? radius * radius : -1.0;
    for (int c = 0; c < 3; ++c) a.origin[c] = tree.origin[c];
    auto out_ptr = [&](void* user, DevTmp& tmp, size_t bytes, void** dev) -> int {
        *dev = nullptr;
        if (!user) return PB200_OK;
        if (!host) { *dev = user; return PB200_OK; }
        PB_CUDA(tmp.alloc(ctx->stream, bytes));
        *dev = tmp.p;
        return PB200_OK;
    };
    void* p = nullptr;
    PB_TRY(out_ptr(idx_out, d_idx, (size_t)nq * k * 4, &p)); a.idx_out = (uint32_t*)p;
    PB_TRY(out_ptr(d2_out, d_d2, (size_t)nq * k * 8, &p)); a.d2_out = (double*)p;
    PB_TRY(out_ptr(counts_out, d_cnt, (size_t)nq * 4, &p)); a.counts_out = (uint32_t*)p;
    PB_TRY(out_ptr(normals_out, d_nrm, (size_t)nq * 24, &p)); a.normals_out = (double*)p;
    PB_TRY(out_ptr(curvature_out, d_curv, (size_t)nq * 8, &p)); a.curvature_out = (double*)p;
    {
        PB_PHASE(ctx, mode == 2 ? "knn.query+normals" : "knn.query");
        if (ctx->knn_heap && mode != 3) PB_TRY(launch_query_heap(ctx, mode, a, ctx->stream));
        else if (k <= 4) launch_query<4>(mode, a, ctx->stream);
        else if (k <= 16) launch_query<16>(mode, a, ctx->stream);
        else if (k <= 32) launch_query<32>(mode, a, ctx->stream);
        else launch_query<64>(mode, a, ctx->stream);
        g_launches++;
    }
    PB_CUDA(cudaGetLastError());
    if (host) {
        if (idx_out) PB_CUDA(cudaMemcpyAsync(idx_out, a.idx_out, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (d2_out) PB_CUDA(cudaMemcpyAsync(d2_out, a.d2_out, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (counts_out) PB_CUDA(cudaMemcpyAsync(counts_out, a.counts_out, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (normals_out) PB_CUDA(cudaMemcpyAsync(normals_out, a.normals_out, (size_t)nq * 24, cudaMemcpyDeviceToHost, ctx->stream));
        if (curvature_out) PB_CUDA(cudaMemcpyAsync(curvature_out, a.curvature_out, (size_t)nq * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(cudaStreamSynchronize(ctx->stream));  // host results must be complete on return
    }
    // the tree and staging buffers are stream-ordered temporaries (DevTmp): released behind the kernels that use them
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_knn_range(pb200_ctx* ctx, const pb200_buffer_desc* buf, uint32_t k, uint64_t first_query, uint64_t n_queries,
                    uint32_t* idx_out, double* d2_out) {
    if (!ctx || !buf || (!idx_out && n_queries)) return set_error(PB200_ERR_INVALID, "null argument");
    int mode = 0;
#ifdef PB200_KNN_DIAGNOSTICS  // diagnostic builds only (make KNN_DIAGNOSTICS=1): idx_out[3*i..] = traversal counters
    if (ctx->knn_stats && k > 4 && k <= 16) mode = 3;
#endif
    return run_query(ctx, buf, mode, k, 0.0, first_query, n_queries, idx_out, d2_out, nullptr, nullptr, nullptr);
}

int pb200_knn(pb200_ctx* ctx, const pb200_buffer_desc* buf, uint32_t k, uint32_t* idx_out, double* d2_out) {
    if (!ctx || !buf || !idx_out) return set_error(PB200_ERR_INVALID, "null argument");
    return pb200_knn_range(ctx, buf, k, 0, buf->len, idx_out, d2_out);
}

int pb200_radius_search(pb200_ctx* ctx, const pb200_buffer_desc* buf, double radius, uint32_t max_neighbors,
                        uint32_t* idx_out, uint32_t* counts_out) {
    if (!ctx || !buf || !idx_out || !counts_out) return set_error(PB200_ERR_INVALID, "null argument");
    if (!(radius >= 0.0)) return set_error(PB200_ERR_INVALID, "radius must be >= 0");
    return run_query(ctx, buf, 1, max_neighbors, radius, 0, buf->len, idx_out, nullptr, counts_out, nullptr, nullptr);
}

int pb200_compute_normals_range(pb200_ctx* ctx, const pb200_buffer_desc* buf, uint32_t k, uint64_t first_query, uint64_t n_queries,
                                double* normals_out, double* curvature_out) {
    if (!ctx || !buf || ((!normals_out || !curvature_out) && n_queries)) return set_error(PB200_ERR_INVALID, "null argument");
    if (buf->len < 3)  // normal_estimation.rs:86-88
        return set_error(PB200_ERR_TOO_FEW_POINTS, "The point cloud is too small. Please use a point cloud that has 3 or more points!");
    if (k < 3)  // :89-91
        return set_error(PB200_ERR_INVALID, "The k nearest neigbors attribute is too small!");
    return run_query(ctx, buf, 2, k, 0.0, first_query, n_queries, nullptr, nullptr, nullptr, normals_out, curvature_out);
}

int pb200_compute_normals(pb200_ctx* ctx, const pb200_buffer_desc* buf, uint32_t k, double* normals_out,
                          double* curvature_out) {
    if (!ctx || !buf || !normals_out || !curvature_out) return set_error(PB200_ERR_INVALID, "null argument");
    return pb200_compute_normals_range(ctx, buf, k, 0, buf->len, normals_out, curvature_out);
}

}  // extern "C"
