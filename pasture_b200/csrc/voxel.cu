// voxel.cu -- voxel-grid downsampling (pasture-algorithms/src/voxel_grid.rs:109-165) as a sort-based GPU pipeline.
//
//   reference                                   here
//   calculate_bounds (:124)                     K5 reduction
//   create_markers_for_axis (:54-79)            host: the same running sum (sequential by definition), uploaded
//   find_leaf per point (:22-51, linear scan)   K7 voxel_key_kernel: O(1) guess + local search on the exact markers,
//                                               nearest-marker fix-up, packed key (ix,iy,iz)
//   sorted Vec<Voxel> with insert (:141-152)    K8 stable LSD radix sort over the significant key bits.  When key bits +
//                                               index bits <= 64 (100 M points at 0.1 m: 34 + 27) the point index rides in
//                                               the low bits of ONE u64 and the sort is keys-only: 16 B/point/pass
//                                               instead of 24, and stability keeps input order inside a voxel for free
//   voxel boundaries                            K8b head count per 2048-point tile -> scan of the tile counts -> emit
//                                               (segment starts, voxel keys, unpacked point indices): 16 B/point read
//   per-voxel attribute reduction (:168-689)    K9 one thread per voxel walks its points IN INPUT ORDER (stable sort),
//                                               so f64 sums round exactly like the reference's sequential loops
// Output order = lexicographic (ix,iy,iz) = ascending packed key.
#include <cub/device/device_radix_sort.cuh>  // fallback for clouds beyond the own sort's 2^30-key limit only

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

struct pb200_voxel_partials {
    pb200_ctx* ctx = nullptr;
    uint64_t len = 0;
    void *keys = nullptr, *counts = nullptr, *sums = nullptr;  // device, from the context's cache
    uint32_t bits_x = 0, bits_y = 0, bits_z = 0;
    uint64_t cells[3] = {0, 0, 0};
    // attribute partials (pb200_voxelgrid_partials_layout): columns per voxel, run lists per "most common value" attribute
    uint32_t n_col = 0, n_modes = 0;
    void* cols = nullptr;
    uint8_t col_is_max[64] = {};
    void* mode_keys[PB200_MAX_ATTRIBUTES] = {};
    void* mode_counts[PB200_MAX_ATTRIBUTES] = {};
    uint64_t mode_len[PB200_MAX_ATTRIBUTES] = {};
};

namespace pb200 {

struct AxisGrid {
    const double2* pairs[3];  // pairs[c][i] = (marker i-1 or 0.0, marker i): both neighbours of a candidate in ONE 16-byte load
    unsigned long long n[3];
    double bmin[3], inv_leaf[3];
    unsigned bits_y, bits_z;
};

__device__ __forceinline__ unsigned long long leaf_index(double p, const double2* __restrict__ m, unsigned long long n,
                                                         double bmin, double inv_leaf) {
    if (n == 0) return 0;  // voxel_grid.rs:31 `!markers.is_empty()`
    // The reference scans for the first index k with !(marker[k] < p) (clamped to the last marker), then steps back to the
    // nearer marker.  Here: guess k from the regular spacing and fetch (marker[k-1], marker[k]) with one 16-byte load; the
    // guess is accepted if the exact (running-sum) markers bracket p, which is the case except for rounding drift, NaN and
    // the clamped ends -- only then walk like the reference.  Lookups of a warp scatter over the tables (x and y of a
    // cloud are unrelated to the point order), so the kernel is bound by the L1 tag rate: one load per axis instead of
    // three is what it is about.
    const double g = (p - bmin) * inv_leaf;
    const int last = (int)n - 1;  // n <= 2^21 + 1
    int k = (g > 0.0) ? ((g < 4.0e6) ? (int)g : last) : 0;
    if (k > last) k = last;
    double2 lh = m[k];  // lo = marker[k-1] (0.0 for k = 0), hi = marker[k]
    if (!((k == 0 || lh.x < p) && (k == last || !(lh.y < p)))) {
        while (k < last && lh.y < p) lh = m[++k];
        while (k > 0 && !(lh.x < p)) lh = m[--k];
    }
    // clamp to the better fitting marker: [k] or [k-1] (voxel_grid.rs:41-49)
    if (k > 0 && __dsub_rn(p, lh.x) < __dsub_rn(lh.y, p)) --k;
    return (unsigned long long)k;
}

// idx_bits > 0: packed mode, keys[i] = voxel key << idx_bits | i (no index array)
__global__ void __launch_bounds__(256) voxel_key_kernel(const uint8_t* __restrict__ pos_base, unsigned long long stride,
                                                        unsigned long long n, AxisGrid g, unsigned idx_bits,
                                                        unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const double* p = reinterpret_cast<const double*>(pos_base + i * stride);
        const unsigned long long ix = leaf_index(p[0], g.pairs[0], g.n[0], g.bmin[0], g.inv_leaf[0]);
        const unsigned long long iy = leaf_index(p[1], g.pairs[1], g.n[1], g.bmin[1], g.inv_leaf[1]);
        const unsigned long long iz = leaf_index(p[2], g.pairs[2], g.n[2], g.bmin[2], g.inv_leaf[2]);
        const unsigned long long key = (((ix << g.bits_y) | iy) << g.bits_z) | iz;
        if (idx_bits) keys[i] = (key << idx_bits) | i;
        else { keys[i] = key; idx[i] = (uint32_t)i; }
    }
}

// ---- voxel boundaries in the sorted key array ---------------------------------------------------------------------
// A tile is HT_ROWS rows of HT_THREADS consecutive elements; thread t owns element r * HT_THREADS + t of every row, so
// all loads are coalesced.  An element is a head if its voxel key differs from its predecessor's.
constexpr int HT_THREADS = 256, HT_ROWS = 8, HT_TILE = HT_THREADS * HT_ROWS;

__device__ __forceinline__ uint32_t tile_heads(const unsigned long long* __restrict__ keys, unsigned long long n, unsigned shift,
                                               unsigned long long base, unsigned long long (&k)[HT_ROWS]) {
    uint32_t flags = 0;
#pragma unroll
    for (int r = 0; r < HT_ROWS; ++r) {
        const unsigned long long i = base + (unsigned long long)r * HT_THREADS + threadIdx.x;
        k[r] = 0;
        if (i < n) {
            k[r] = keys[i];
            const bool head = i == 0 || (keys[i - 1] >> shift) != (k[r] >> shift);
            flags |= head ? (1u << r) : 0u;
        }
    }
    return flags;
}

__global__ void __launch_bounds__(HT_THREADS) heads_count_kernel(const unsigned long long* __restrict__ keys, unsigned long long n,
                                                                 unsigned shift, uint32_t* __restrict__ tile_counts) {
    __shared__ uint32_t warp_sum[HT_THREADS / 32];
    unsigned long long k[HT_ROWS];
    uint32_t c = __popc(tile_heads(keys, n, shift, (unsigned long long)blockIdx.x * HT_TILE, k));
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < HT_THREADS / 32; ++w) t += warp_sum[w];
        tile_counts[blockIdx.x] = t;
    }
}

// starts[rank] = sorted position of the rank-th voxel's first point, voxel_keys[rank] = its key; in packed mode the point
// indices are unpacked into idx_out on the way (the keys are being read anyway)
__global__ void __launch_bounds__(HT_THREADS) heads_emit_kernel(const unsigned long long* __restrict__ keys, unsigned long long n,
                                                                unsigned shift, const uint32_t* __restrict__ tile_offsets,
                                                                uint32_t n_voxels, uint32_t* __restrict__ starts,
                                                                unsigned long long* __restrict__ voxel_keys,
                                                                uint32_t* __restrict__ idx_out) {
    __shared__ uint32_t part[HT_ROWS][HT_THREADS / 32];  // heads per (row, warp), then exclusive prefix in row-major order
    unsigned long long k[HT_ROWS];
    const unsigned long long base = (unsigned long long)blockIdx.x * HT_TILE;
    const uint32_t flags = tile_heads(keys, n, shift, base, k);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t lane_rank[HT_ROWS];
#pragma unroll
    for (int r = 0; r < HT_ROWS; ++r) {
        const uint32_t b = __ballot_sync(0xffffffffu, (flags >> r) & 1u);
        lane_rank[r] = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) part[r][warp] = __popc(b);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 64 (row, warp) counts in row-major order by one warp: two entries per lane
        static_assert(HT_ROWS * (HT_THREADS / 32) == 64, "two entries per lane");
        uint32_t* flat = &part[0][0];
        const uint32_t c0 = flat[lane], c1 = flat[lane + 32];
        uint32_t x0 = c0, x1 = c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
            if ((int)lane >= o) { x0 += y0; x1 += y1; }
        }
        const uint32_t first_half = __shfl_sync(0xffffffffu, x0, 31), base_off = tile_offsets[blockIdx.x];
        flat[lane] = base_off + x0 - c0;
        flat[lane + 32] = base_off + first_half + x1 - c1;
    }
    __syncthreads();
    const unsigned long long idx_mask = shift ? ((1ull << shift) - 1ull) : 0ull;
#pragma unroll
    for (int r = 0; r < HT_ROWS; ++r) {
        const unsigned long long i = base + (unsigned long long)r * HT_THREADS + threadIdx.x;
        if (i >= n) continue;
        if (idx_out) idx_out[i] = (uint32_t)(k[r] & idx_mask);
        if ((flags >> r) & 1u) {
            const uint32_t rank = part[r][warp] + lane_rank[r];
            starts[rank] = (uint32_t)i;
            voxel_keys[rank] = k[r] >> shift;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) starts[n_voxels] = (uint32_t)n;
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

__device__ __forceinline__ double ld_gather_f64(const uint8_t* p) {
    double v;
    asm volatile("ld.global.nc.L2::64B.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// three consecutive doubles at an 8-byte aligned address as one 16-byte and one 8-byte load
__device__ __forceinline__ void ld_gather_pos24(const uint8_t* p, double& x, double& y, double& z) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        asm volatile("ld.global.nc.L2::64B.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p));
        z = ld_gather_f64(p + 16);
    } else {
        x = ld_gather_f64(p);
        asm volatile("ld.global.nc.L2::64B.v2.f64 {%0, %1}, [%2];" : "=d"(y), "=d"(z) : "l"(p + 8));
    }
}

// ---- K8b + K9 fused: voxel boundaries AND the per-voxel position reduction in one pass over the sorted keys ----------
// One CTA per tile of ER_TILE sorted keys (same striped ownership as the heads kernels).  Every THREAD gathers the
// positions of its own ER_ROWS points -- ER_GB independent 24-byte gathers in flight per thread, all 256 threads busy: the
// memory-level parallelism the thread-per-voxel kernel lacked (it walked its points one dependent gather at a time and sat
// at 6 % issue utilisation) -- and parks them in shared memory.  Then the thread that owns a voxel's FIRST point adds the
// voxel's points up from shared memory in sorted = input order, i.e. with exactly the roundings of the reference's
// sequential loop (voxel_grid.rs:339-386).  A voxel that runs past the tile end is finished by its owner straight from
// global memory (one thread per tile at most).  Writes centroids (OUT = 0) or sums + counts (OUT = 1, shard partials),
// the voxel keys, and -- only when other attributes need them (EMIT_INDEX) -- segment starts and unpacked indices.
constexpr int ER_THREADS = HT_THREADS, ER_ROWS = HT_ROWS, ER_TILE = HT_TILE, ER_GB = 4;
constexpr size_t ER_SMEM = (size_t)ER_TILE * 24 + 64 * 4;

template <int OUT, bool EMIT_INDEX>
__global__ void __launch_bounds__(ER_THREADS, 3)
voxel_emit_reduce_kernel(const unsigned long long* __restrict__ keys, unsigned long long n, unsigned shift,
                         const uint32_t* __restrict__ tile_offsets, uint32_t n_voxels, const uint32_t* __restrict__ idx_in,
                         const uint8_t* __restrict__ pos, unsigned long long pstride, uint32_t* __restrict__ starts, unsigned long long* __restrict__ voxel_keys, uint32_t* __restrict__ idx_out,
                         double* __restrict__ out3, uint32_t* __restrict__ counts) {
    extern __shared__ __align__(16) uint8_t er_smem[];
    double* sx = reinterpret_cast<double*>(er_smem);
    double* sy = sx + ER_TILE;
    double* sz = sy + ER_TILE;
    uint32_t* s_flags = reinterpret_cast<uint32_t*>(sz + ER_TILE);  // [64]: bit l of word r * 8 + w <-> element r * 256 + w * 32 + l
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long base = (unsigned long long)blockIdx.x * ER_TILE;
    const uint32_t npts = (uint32_t)((n - base) < (unsigned long long)ER_TILE ? (n - base) : (unsigned long long)ER_TILE);
    const unsigned long long idx_mask = shift ? ((1ull << shift) - 1ull) : 0ull;
    unsigned long long k[ER_ROWS];
    const uint32_t flags = tile_heads(keys, n, shift, base, k);
    // per (row, warp) head counts: lane (r * 8 + w) & 31 of EVERY warp holds entry r * 8 + w (entries 0..31 in c0, 32..63 in c1)
    uint32_t lane_rank[ER_ROWS];
#pragma unroll
    for (int r = 0; r < ER_ROWS; ++r) {
        const uint32_t b = __ballot_sync(0xffffffffu, (flags >> r) & 1u);
        lane_rank[r] = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) s_flags[r * 8 + warp] = b;
    }
    // gather this thread's points, ER_GB at a time, into shared memory (element e = r * 256 + t: conflict-free)
#pragma unroll
    for (int r0 = 0; r0 < ER_ROWS; r0 += ER_GB) {
        double x[ER_GB], y[ER_GB], z[ER_GB];
#pragma unroll
        for (int g = 0; g < ER_GB; ++g) {
            const int r = r0 + g;
            const uint32_t e = (uint32_t)r * ER_THREADS + threadIdx.x;
            x[g] = y[g] = z[g] = 0.0;
            if (e < npts) {
                const uint32_t j = shift ? (uint32_t)(k[r] & idx_mask) : idx_in[base + e];
                if (EMIT_INDEX && shift) idx_out[base + e] = j;
                ld_gather_pos24(pos + (unsigned long long)j * pstride, x[g], y[g], z[g]);
            }
        }
#pragma unroll
        for (int g = 0; g < ER_GB; ++g) {
            const uint32_t e = (uint32_t)(r0 + g) * ER_THREADS + threadIdx.x;
            sx[e] = x[g]; sy[e] = y[g]; sz[e] = z[g];
        }
    }
    __syncthreads();
    // exclusive scan of the 64 (row, warp) head counts, done by every warp for itself (two entries per lane)
    const uint32_t c0 = __popc(s_flags[lane]), c1 = __popc(s_flags[lane + 32]);
    uint32_t x0 = c0, x1 = c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
        if ((int)lane >= o) { x0 += y0; x1 += y1; }
    }
    const uint32_t first_half = __shfl_sync(0xffffffffu, x0, 31), tile_off = tile_offsets[blockIdx.x];
    const uint32_t ex0 = tile_off + x0 - c0, ex1 = tile_off + first_half + x1 - c1;
#pragma unroll
    for (int r = 0; r < ER_ROWS; ++r) {
        const uint32_t entry = (uint32_t)r * 8u + warp;  // warp-uniform
        const uint32_t row_base = __shfl_sync(0xffffffffu, entry < 32u ? ex0 : ex1, entry & 31u);
        if (!((flags >> r) & 1u)) continue;
        const uint32_t e = (uint32_t)r * ER_THREADS + threadIdx.x;
        // end of this voxel inside the tile: the next head bit after e, else the tile end
        uint32_t end = npts;
        {
            const uint32_t j = e + 1u;
            if (j < (uint32_t)ER_TILE) {
                uint32_t w = j >> 5, m = s_flags[w] >> (j & 31u);
                if (m) end = j + (uint32_t)__ffs((int)m) - 1u;
                else {
                    for (++w; w < 64u; ++w) {
                        m = s_flags[w];
                        if (m) { end = w * 32u + (uint32_t)__ffs((int)m) - 1u; break; }
                    }
                }
            }
            end = end < npts ? end : npts;
        }
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (uint32_t j = e; j < end; ++j) {
            ax = __dadd_rn(ax, sx[j]);
            ay = __dadd_rn(ay, sy[j]);
            az = __dadd_rn(az, sz[j]);
        }
        unsigned long long cnt = end - e;
        const unsigned long long vkey = k[r] >> shift;
        if (end == npts) {  // the voxel may continue in the following tiles: finish it from global memory
            for (unsigned long long g = base + npts; g < n; ++g, ++cnt) {
                const unsigned long long kk = keys[g];
                if ((kk >> shift) != vkey) break;
                const uint32_t j = shift ? (uint32_t)(kk & idx_mask) : idx_in[g];
                double px, py, pz;
                ld_gather_pos24(pos + (unsigned long long)j * pstride, px, py, pz);
                ax = __dadd_rn(ax, px);
                ay = __dadd_rn(ay, py);
                az = __dadd_rn(az, pz);
            }
        }
        const uint32_t rank = row_base + lane_rank[r];
        voxel_keys[rank] = vkey;
        if (EMIT_INDEX) starts[rank] = (uint32_t)(base + e);
        if (OUT == 0) {  // voxel_grid.rs:382-386
            const double c = (double)cnt;
            out3[3 * (size_t)rank] = ax / c; out3[3 * (size_t)rank + 1] = ay / c; out3[3 * (size_t)rank + 2] = az / c;
        } else if (OUT == 1) {
            out3[3 * (size_t)rank] = ax; out3[3 * (size_t)rank + 1] = ay; out3[3 * (size_t)rank + 2] = az;
            counts[rank] = (uint32_t)cnt;
        }
    }
    if (EMIT_INDEX && blockIdx.x == 0 && threadIdx.x == 0) starts[n_voxels] = (uint32_t)n;
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
    // sharded grid: instead of the finished value, the raw f64 component sums (mean kinds) / running maximum
    // (max-pool kinds) of the voxel go to partial[v * partial_stride ...] -- the division and the cast happen after the merge
    double* partial;
    uint32_t partial_stride;
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
            if constexpr (sizeof(S) == 8 && KIND == R_MEAN_VEC_F64) {
                if (al) {  // random 24-byte gathers: 64-byte L2 fills, and TWO loads per point (16 + 8 bytes, whichever
                           // half is 16-byte aligned): the L1 tag stage, not DRAM, limits this kernel once fills are small
                    double x, y, z;
                    ld_gather_pos24(p, x, y, z);
                    sx = __dadd_rn(sx, x);
                    sy = __dadd_rn(sy, y);
                    sz = __dadd_rn(sz, z);
                    continue;
                }
            }
            sx = __dadd_rn(sx, (double)ld_attr<S>(p, al));
            sy = __dadd_rn(sy, (double)ld_attr<S>(p + sizeof(S), al));
            sz = __dadd_rn(sz, (double)ld_attr<S>(p + 2 * sizeof(S), al));
        }
        if (a.partial) { double* q = a.partial + v * a.partial_stride; q[0] = sx; q[1] = sy; q[2] = sz; return; }
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
        if (a.partial) { a.partial[v * a.partial_stride] = s; return; }
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
        if (a.partial) { a.partial[v * a.partial_stride] = cur; return; }
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
__global__ void __launch_bounds__(256) mode_keys_kernel(const uint32_t* __restrict__ starts, uint32_t n_voxels,
                                                        const uint32_t* __restrict__ sorted_idx, unsigned long long n,
                                                        const uint8_t* __restrict__ src, unsigned long long stride, int aligned,
                                                        unsigned long long* __restrict__ keys) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        uint32_t lo = 0, hi = n_voxels;  // voxel of sorted position i: last v with starts[v] <= i
        while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (starts[mid] <= i) lo = mid; else hi = mid; }
        const unsigned long long voxel = lo;
        const long long val = (long long)ld_attr<S>(src + (unsigned long long)sorted_idx[i] * stride, aligned != 0);
        constexpr long long bias = (S(-1) < S(0)) ? 32768 : 0;  // keeps signed values ordered as unsigned 16-bit fields
        keys[i] = (voxel << 16) | (unsigned long long)((val + bias) & 0xFFFF);
    }
}

// one vote per run of equal (voxel, value) keys: the thread at a run's head finds its end by binary search in the sorted
// keys (runs are long in the crowded case this path exists for), then length << 16 | 65535 - value competes per voxel
__global__ void __launch_bounds__(256) run_vote_kernel(const unsigned long long* __restrict__ keys, unsigned long long n,
                                                       unsigned long long* __restrict__ best) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const unsigned long long k = keys[i];
        if (i > 0 && keys[i - 1] == k) continue;  // not a head
        unsigned long long lo = i, hi = n;        // last index with keys[idx] == k lies in [lo, hi)
        while (hi - lo > 1) { const unsigned long long mid = lo + (hi - lo) / 2; if (keys[mid] == k) lo = mid; else hi = mid; }
        const unsigned long long len = lo - i + 1ull;
        atomicMax(&best[k >> 16], (len << 16) | (65535ull - (k & 0xFFFFull)));
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

// ---- sharded voxel grid: partial sums per shard, merge of exchanged partials ------------------------
__global__ void __launch_bounds__(256) iota_kernel(uint32_t* __restrict__ out, unsigned long long n) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) out[i] = (uint32_t)i;
}

// per merged voxel: add up the partials in the order they were concatenated (= source rank order: the sort is stable)
__global__ void __launch_bounds__(128) partials_merge_kernel(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ sorted_idx,
                                                             unsigned long long n_voxels, const uint32_t* __restrict__ in_counts,
                                                             const double* __restrict__ in_sums, uint32_t* __restrict__ counts,
                                                             double* __restrict__ sums) {
    const unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_voxels) return;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    uint32_t c = 0;
    for (uint32_t k = starts[v], e = starts[v + 1]; k < e; ++k) {
        const uint32_t j = sorted_idx[k];
        c += in_counts[j];
        sx = __dadd_rn(sx, in_sums[3 * (size_t)j]); sy = __dadd_rn(sy, in_sums[3 * (size_t)j + 1]); sz = __dadd_rn(sz, in_sums[3 * (size_t)j + 2]);
    }
    counts[v] = c;
    sums[3 * v] = sx; sums[3 * v + 1] = sy; sums[3 * v + 2] = sz;
}

// centroid = sum / count (voxel_grid.rs:382-386)
__global__ void __launch_bounds__(256) partials_centroid_kernel(const uint32_t* __restrict__ counts, const double* __restrict__ sums,
                                                                unsigned long long n_voxels, double* __restrict__ out) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_voxels; v += step) {
        const double c = (double)counts[v];
        out[3 * v] = sums[3 * v] / c; out[3 * v + 1] = sums[3 * v + 1] / c; out[3 * v + 2] = sums[3 * v + 2] / c;
    }
}

// Everything between the bounds and the per-voxel reductions: markers on the host (voxel_grid.rs:63-77), keys, sort,
// voxel boundaries.  With the AABB of the WHOLE cloud every shard of a sharded cloud derives the same markers and
// therefore the same global voxel keys.
// ---- sharded grid, attributes other than POSITION_3D (SURVEY 8e "per-attribute partials") ---------------------------
// Per voxel and shard the mean / max-pool attributes travel as 8-byte COLUMNS (component sums, running maxima), the
// "most common value" attributes as RUN LISTS (composite key = voxel key << 16 | biased value, count).
struct ColOps { uint8_t is_max[64]; };  // per column: 0 = partial sums are added, 1 = the maximum is taken

// merge of concatenated partials, generalised: counts and position sums as before, plus n_col columns
__global__ void __launch_bounds__(128) partials_merge_columns_kernel(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ sorted_idx,
                                                                     unsigned long long n_voxels, const double* __restrict__ in_cols,
                                                                     uint32_t n_col, ColOps ops, double* __restrict__ out_cols) {
    const unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_voxels) return;
    for (uint32_t c = 0; c < n_col; ++c) {
        double acc = 0.0;  // sums start at 0; the max-pool folds start at 0.0 as well (voxel_grid.rs:168-215)
        for (uint32_t k = starts[v], e = starts[v + 1]; k < e; ++k) {
            const double x = in_cols[(size_t)sorted_idx[k] * n_col + c];
            if (ops.is_max[c]) { if (x > acc) acc = x; }
            else acc = __dadd_rn(acc, x);
        }
        out_cols[v * n_col + c] = acc;
    }
}

// runs of equal (voxel rank << 16 | biased value) keys of one shard -> (voxel KEY << 16 | biased value, run length)
__global__ void __launch_bounds__(256) mode_runs_kernel(const unsigned long long* __restrict__ run_keys, const uint32_t* __restrict__ run_starts,
                                                        unsigned long long n_runs, const unsigned long long* __restrict__ voxel_keys,
                                                        unsigned long long* __restrict__ out_keys, uint32_t* __restrict__ out_counts) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += step) {
        const unsigned long long k = run_keys[r];
        out_keys[r] = (voxel_keys[k >> 16] << 16) | (k & 0xFFFFull);
        out_counts[r] = run_starts[r + 1] - run_starts[r];
    }
}

// merged runs vote: total count of a (voxel, value) pair over all shards competes for its voxel (found by binary search
// in the merged voxel keys); the largest count wins, ties go to the smallest value -- the rule of the single-device filter
__global__ void __launch_bounds__(256) mode_merge_vote_kernel(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ sorted_idx,
                                                              const unsigned long long* __restrict__ run_keys, unsigned long long n_runs,
                                                              const uint32_t* __restrict__ in_counts,
                                                              const unsigned long long* __restrict__ voxel_keys, unsigned long long n_voxels,
                                                              unsigned long long* __restrict__ best) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += step) {
        unsigned long long total = 0;
        for (uint32_t k = starts[r], e = starts[r + 1]; k < e; ++k) total += in_counts[sorted_idx[k]];
        const unsigned long long comp = run_keys[r], vkey = comp >> 16;
        unsigned long long lo = 0, hi = n_voxels;  // first voxel with key >= vkey (it exists: every run belongs to a merged voxel)
        while (lo < hi) { const unsigned long long mid = lo + (hi - lo) / 2; if (voxel_keys[mid] < vkey) lo = mid + 1; else hi = mid; }
        if (lo < n_voxels && voxel_keys[lo] == vkey) atomicMax(&best[lo], (total << 16) | (65535ull - (comp & 0xFFFFull)));
    }
}

// merged columns -> attribute values: the division by the point count and the reference's casts (voxel_grid.rs:461-679)
__global__ void __launch_bounds__(256) finalize_columns_kernel(int kind, const double* __restrict__ cols, uint32_t n_col, uint32_t col0,
                                                               const uint32_t* __restrict__ counts, unsigned long long n_voxels,
                                                               uint8_t* __restrict__ dst, uint32_t dst_size) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_voxels; v += step) {
        const double* c = cols + v * n_col + col0;
        uint8_t* out = dst + v * dst_size;
        const double cnt = (double)counts[v];
        auto u16 = [](double x) { unsigned long long u = f64_as_u64_sat(x); return (uint16_t)(u > 65535ull ? 65535ull : u); };
        switch (kind) {
            case R_MEAN_U16: { const uint16_t r = u16(c[0] / cnt); memcpy(out, &r, 2); break; }
            case R_MEAN_VEC_U16: { const uint16_t r[3] = {u16(c[0] / cnt), u16(c[1] / cnt), u16(c[2] / cnt)}; memcpy(out, r, 6); break; }
            case R_MEAN_VEC_F32: { const float r[3] = {(float)(c[0] / cnt), (float)(c[1] / cnt), (float)(c[2] / cnt)}; memcpy(out, r, 12); break; }
            case R_MAX_U8: { unsigned long long u = f64_as_u64_sat(c[0]); const uint8_t r = (uint8_t)(u > 255ull ? 255ull : u); memcpy(out, &r, 1); break; }
            case R_MAX_F64: { memcpy(out, c, 8); break; }
            case R_MAX_U64: { const unsigned long long r = f64_as_u64_sat(c[0]); memcpy(out, &r, 8); break; }
            default: break;
        }
    }
}

struct VoxelIndex {
    DevTmp sorted_keys, sorted_idx, starts, voxel_keys, tiles;  // N, N, V+1, V, tiles+1
    uint64_t V = 0, n = 0;
    unsigned bits_x = 0, bits_y = 0, bits_z = 0, shift = 0;
    bool packed = false;
    uint32_t n_tiles = 0;
    uint64_t cells[3] = {0, 0, 0};
    const uint8_t* ppos = nullptr;  // positions the sorted indices refer to
    uint64_t pstride = 0;
};

// Step 1: markers on the host (voxel_grid.rs:63-77), keys, sort, heads per tile + scan -> V is known on the host.
static int voxel_index_sort(pb200_ctx* ctx, const uint8_t* ppos, uint64_t pstride, uint64_t n, const double bmin[3],
                            const double bmax[3], const double leaf[3], VoxelIndex* vi) {
    cudaStream_t st = ctx->stream;
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
    vi->bits_x = bits_x; vi->bits_y = bits_y; vi->bits_z = bits_z;
    vi->n = n; vi->ppos = ppos; vi->pstride = pstride;
    for (int c = 0; c < 3; ++c) vi->cells[c] = markers[c].size();
    DevTmp d_markers;
    const size_t nm_total = markers[0].size() + markers[1].size() + markers[2].size();
    PB_CUDA(d_markers.alloc(st, (nm_total + 1) * sizeof(double2)));
    AxisGrid grid;
    {
        PB_PHASE(ctx, "host.markers_upload");
        // (marker[i-1], marker[i]) pairs, staged in pinned memory when they fit (a pageable source would be copied
        // synchronously through the driver's own staging buffer)
        std::vector<double2> pairs_host;
        double2* hp = nullptr;
        if ((nm_total + 1) * sizeof(double2) <= ctx->h_stage_bytes) hp = (double2*)ctx->h_stage;
        else { pairs_host.resize(nm_total + 1); hp = pairs_host.data(); }
        size_t off = 0;
        for (int c = 0; c < 3; ++c) {
            grid.pairs[c] = (const double2*)d_markers.p + off;
            grid.n[c] = markers[c].size();
            grid.bmin[c] = bmin[c];
            grid.inv_leaf[c] = 1.0 / leaf[c];
            for (size_t i = 0; i < markers[c].size(); ++i) hp[off + i] = make_double2(i ? markers[c][i - 1] : 0.0, markers[c][i]);
            off += markers[c].size();
        }
        if (nm_total) {
            // (h_stage is free: every call that uploads from it synchronises on its voxel count further down)
            PB_CUDA(cudaMemcpyAsync(d_markers.p, hp, nm_total * sizeof(double2), cudaMemcpyHostToDevice, st));
            if (hp != (double2*)ctx->h_stage) PB_CUDA(cudaStreamSynchronize(st));  // pairs_host dies with this scope
        }
        grid.bits_y = bits_y;
        grid.bits_z = bits_z;
    }

    // ---- keys, sort, segments ---------------------------------------------------------------------------------
    // packed mode: the point index rides in the low bits of the key, the sort is keys-only
    const unsigned key_bits = bits_x + bits_y + bits_z, idx_bits_needed = bits_for(n);
    const bool packed = key_bits + idx_bits_needed <= 64;
    const unsigned shift = packed ? idx_bits_needed : 0;
    vi->packed = packed; vi->shift = shift;
    DevTmp d_keys, d_idx, d_tmp;
    DevTmp &d_keys2 = vi->sorted_keys, &d_idx2 = vi->sorted_idx, &d_tiles = vi->tiles;
    {
        PB_PHASE(ctx, "host.alloc_keys");
        PB_CUDA(d_keys.alloc(st, n * 8)); PB_CUDA(d_keys2.alloc(st, n * 8));
        if (!packed) { PB_CUDA(d_idx.alloc(st, n * 4)); PB_CUDA(d_idx2.alloc(st, n * 4)); }
    }
    if (((uintptr_t)ppos & 7) || (pstride & 7)) return set_error(PB200_ERR_UNSUPPORTED, "POSITION_3D must be 8-byte aligned in memory");
    const unsigned long long cap = (unsigned long long)ctx->sm_count * 16;
    const unsigned blocks = (unsigned)(((n + 255) / 256) < cap ? ((n + 255) / 256) : cap);
    {
        PB_PHASE(ctx, "voxel.keys");
        voxel_key_kernel<<<blocks, 256, 0, st>>>(ppos, pstride, n, grid, shift, (unsigned long long*)d_keys.p, (uint32_t*)d_idx.p);
        g_launches++;
    }
    // K8: own one-sweep radix sort (radix_sort.cu); cub only for clouds beyond its 2^30-key limit
    if (n < (1ull << 30)) {
        bool in_alt = false;
        PB_TRY(radix_sort_u64(ctx, (unsigned long long*)d_keys.p, (unsigned long long*)d_keys2.p, packed ? nullptr : (uint32_t*)d_idx.p,
                              packed ? nullptr : (uint32_t*)d_idx2.p, n, (int)shift, (int)(shift + key_bits), &in_alt));
        if (!in_alt) {  // the result must end up in d_keys2 / d_idx2 (= vi->sorted_keys / sorted_idx)
            std::swap(d_keys.p, d_keys2.p);
            if (!packed) std::swap(d_idx.p, d_idx2.p);
        }
    } else {
        size_t tmp_bytes = 0;
        if (packed) {
            cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (const unsigned long long*)d_keys.p, (unsigned long long*)d_keys2.p, (int)n,
                                           (int)shift, (int)(shift + key_bits), st);
            PB_CUDA(d_tmp.alloc(st, tmp_bytes));
            PB_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp.p, tmp_bytes, (const unsigned long long*)d_keys.p, (unsigned long long*)d_keys2.p, (int)n,
                                                   (int)shift, (int)(shift + key_bits), st));
        } else {
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned long long*)d_keys.p, (unsigned long long*)d_keys2.p,
                                            (const uint32_t*)d_idx.p, (uint32_t*)d_idx2.p, (int)n, 0, (int)key_bits, st);
            PB_CUDA(d_tmp.alloc(st, tmp_bytes));
            PB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, (const unsigned long long*)d_keys.p, (unsigned long long*)d_keys2.p,
                                                    (const uint32_t*)d_idx.p, (uint32_t*)d_idx2.p, (int)n, 0, (int)key_bits, st));
        }
        g_launches += (uint64_t)((key_bits + 7) / 8) + 1;
    }
    // voxel boundaries: heads per tile -> scan (the emit step follows once V is known on the host)
    const uint32_t n_tiles = (uint32_t)((n + HT_TILE - 1) / HT_TILE);
    vi->n_tiles = n_tiles;
    PB_CUDA(d_tiles.alloc(st, ((size_t)n_tiles + 1) * 4));
    uint32_t* d_total = (uint32_t*)d_tiles.p + n_tiles;
    {
        PB_PHASE(ctx, "voxel.heads_count+scan");
        heads_count_kernel<<<n_tiles, HT_THREADS, 0, st>>>((const unsigned long long*)d_keys2.p, n, shift, (uint32_t*)d_tiles.p);
        PB_TRY(exclusive_scan_u32(ctx, (uint32_t*)d_tiles.p, n_tiles, d_total));
        g_launches += 2;
    }
    uint32_t* h_total = (uint32_t*)ctx->h_scratch;
    {
        PB_PHASE(ctx, "host.sync_voxel_count");
        PB_CUDA(cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
    }
    vi->V = *h_total;
    return PB200_OK;
}

// Step 2: voxel keys (always), the per-voxel position reduction fused into the same pass (out_kind 0: centroids into
// out3 = V x 3 doubles; 1: sums into out3 + counts; 2: none), and -- if other attributes are reduced afterwards
// (need_index) -- segment starts and the unpacked sorted point indices.
static int voxel_index_emit(pb200_ctx* ctx, VoxelIndex* vi, int out_kind, double* out3, uint32_t* counts, bool need_index) {
    cudaStream_t st = ctx->stream;
    const uint64_t V = vi->V, n = vi->n;
    if (out_kind == 2) need_index = true;  // boundaries only: the caller wants starts / indices, nothing else comes out
    PB_CUDA(vi->voxel_keys.alloc(st, V * 8 + 8));
    if (need_index) {
        PB_CUDA(vi->starts.alloc(st, (V + 1) * 4));
        if (vi->packed) PB_CUDA(vi->sorted_idx.alloc(st, n * 4));
    }
    const unsigned long long* keys = (const unsigned long long*)vi->sorted_keys.p;
    const uint32_t* tiles = (const uint32_t*)vi->tiles.p;
    uint32_t* starts = (uint32_t*)vi->starts.p;
    unsigned long long* vkeys = (unsigned long long*)vi->voxel_keys.p;
    uint32_t* idx_out = vi->packed ? (uint32_t*)vi->sorted_idx.p : nullptr;
    const uint32_t* idx_in = vi->packed ? nullptr : (const uint32_t*)vi->sorted_idx.p;
    if (out_kind == 2) {
        PB_PHASE(ctx, "voxel.heads_emit");
        heads_emit_kernel<<<vi->n_tiles, HT_THREADS, 0, st>>>(keys, n, vi->shift, tiles, (uint32_t)V, starts, vkeys, need_index ? idx_out : nullptr);
        g_launches++;
        PB_CUDA(cudaGetLastError());
        return PB200_OK;
    }
    if (!ctx->voxel_attr_set) {
        PB_CUDA(cudaFuncSetAttribute(voxel_emit_reduce_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ER_SMEM));
        PB_CUDA(cudaFuncSetAttribute(voxel_emit_reduce_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ER_SMEM));
        PB_CUDA(cudaFuncSetAttribute(voxel_emit_reduce_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ER_SMEM));
        PB_CUDA(cudaFuncSetAttribute(voxel_emit_reduce_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ER_SMEM));
        ctx->voxel_attr_set = true;
    }
    PB_PHASE(ctx, "voxel.emit+reduce");
#define PB_ER_LAUNCH(O, E)                                                                                                        \
    voxel_emit_reduce_kernel<O, E><<<vi->n_tiles, ER_THREADS, ER_SMEM, st>>>(keys, n, vi->shift, tiles, (uint32_t)V, idx_in, vi->ppos, \
                                                                            vi->pstride, starts, vkeys, idx_out, out3, counts)
    if (out_kind == 0) { if (need_index) PB_ER_LAUNCH(0, true); else PB_ER_LAUNCH(0, false); }
    else { if (need_index) PB_ER_LAUNCH(1, true); else PB_ER_LAUNCH(1, false); }
#undef PB_ER_LAUNCH
    g_launches++;
    PB_CUDA(cudaGetLastError());
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_voxelgrid_filter(pb200_ctx* ctx, const pb200_buffer_desc* src, double lx, double ly, double lz,
                           const pb200_layout* dst_layout, int32_t dst_kind, int32_t dst_memspace,
                           pb200_result_buffer** out) {
    if (!ctx || !dst_layout || !out) return set_error(PB200_ERR_INVALID, "null argument");
    *out = nullptr;
    PB_TRY(validate_desc(src, "source buffer"));
    PB_DEVICE(ctx);
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
    PB_PHASE(ctx, "voxel.call");

    // ---- source attribute streams on the device --------------------------------------------------------------
    std::vector<DevTmp> staged(src->layout->attrs.size() + 1);
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
                PB_CUDA(staged.back().alloc(st, (size_t)(n * src->layout->size)));
                PB_CUDA(cudaMemcpyAsync(staged.back().p, src->aos, (size_t)(n * src->layout->size), cudaMemcpyHostToDevice, st));
                d_aos = (const uint8_t*)staged.back().p;
            }
        } else if (!d_cols[(size_t)idx]) {
            const size_t bytes = (size_t)(n * src->layout->attrs[(size_t)idx].size);
            PB_CUDA(staged[(size_t)idx].alloc(st, bytes));
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
    { PB_PHASE(ctx, "voxel.bounds"); PB_TRY(pb200_calculate_bounds(ctx, &dsrc, bmin, bmax, &some)); }
    if (!some) return set_error(PB200_ERR_INVALID, "calculate_bounds returned None");
    uint64_t pstride = 0;
    const uint8_t* ppos = attr_ptr(pi, &pstride);
    // the Position3D centroid is reduced inside the boundary pass (voxel_emit_reduce_kernel); segment starts and sorted
    // indices are only materialised when other attributes are reduced afterwards
    int a_pos = -1;
    for (size_t a = 0; a < rules.size(); ++a)
        if (rules[a]->kind == R_MEAN_VEC_F64) a_pos = (int)a;
    const bool need_index = rules.size() > (a_pos >= 0 ? 1u : 0u);
    VoxelIndex vi;
    PB_TRY(voxel_index_sort(ctx, ppos, pstride, n, bmin, bmax, leaf, &vi));
    const uint64_t V = vi.V;
    const unsigned bits_y = vi.bits_y, bits_z = vi.bits_z;
    pb200_result_buffer* res = new pb200_result_buffer();
    res->ctx = ctx;
    res->layout = *dst_layout;
    res->kind = dst_kind;
    res->memspace = dst_memspace;
    res->len = V;
    std::vector<void*> d_out(dst_layout->attrs.size(), nullptr);
    auto fail = [&](int rc) {
        for (void* p : d_out) if (p) cache_free(ctx, p);
        if (res->d_packed_keys) cache_free(ctx, res->d_packed_keys);
        delete res;
        return rc;
    };
    {
        PB_PHASE(ctx, "host.alloc_result");
        for (size_t a = 0; a < dst_layout->attrs.size(); ++a) {
            cudaError_t e = cache_alloc(ctx, &d_out[a], (size_t)(V * dst_layout->attrs[a].size) + 16);
            if (e != cudaSuccess) return fail(cuda_error(e, "result column allocation"));
        }
    }
    {
        int rc = voxel_index_emit(ctx, &vi, a_pos >= 0 ? 0 : 2, a_pos >= 0 ? (double*)d_out[(size_t)a_pos] : nullptr, nullptr, need_index);
        if (rc < 0) return fail(rc);
    }
    DevTmp &d_idx2 = vi.sorted_idx, &d_starts = vi.starts, &d_vkeys = vi.voxel_keys;
    const unsigned long long cap = (unsigned long long)ctx->sm_count * 16;
    const unsigned blocks = (unsigned)(((n + 255) / 256) < cap ? ((n + 255) / 256) : cap);
    uint32_t* h_total = (uint32_t*)ctx->h_scratch;

    // the most crowded voxel decides how the mode attributes are reduced
    bool any_mode = false;
    for (const Rule* r : rules) any_mode = any_mode || r->kind == R_MODE || r->kind == R_MODE_BOOL;
    uint32_t max_occ = 0;
    DevTmp d_mode_keys, d_mode_keys2, d_best, d_occ;
    if (any_mode) {
        if (d_occ.alloc(st, 4) != cudaSuccess) return fail(set_error(PB200_ERR_OOM, "out of device memory"));
        cudaMemsetAsync(d_occ.p, 0, 4, st);
        max_occupancy_kernel<<<blocks, 256, 0, st>>>((const uint32_t*)d_starts.p, V, (uint32_t*)d_occ.p);
        g_launches++;
        cudaMemcpyAsync(h_total, d_occ.p, 4, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) return fail(cuda_error(cudaGetLastError(), "max occupancy"));
        max_occ = *h_total;
        if (max_occ > MODE_THREAD_LIMIT) {
            if (d_mode_keys.alloc(st, n * 8) != cudaSuccess || d_mode_keys2.alloc(st, n * 8) != cudaSuccess ||
                d_best.alloc(st, (V + 1) * 8) != cudaSuccess)
                return fail(set_error(PB200_ERR_OOM, "out of device memory"));
        }
    }

    // ---- per-attribute reductions into a columnar staging result -----------------------------------------------
    for (size_t a = 0; a < dst_layout->attrs.size(); ++a) {
        if ((int)a == a_pos) continue;  // done by the fused boundary pass
        ReduceArgs ra;
        ra.partial = nullptr;
        ra.partial_stride = 0;
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
            const uint32_t *sts = (const uint32_t*)d_starts.p, *si2 = (const uint32_t*)d_idx2.p;
            const uint32_t nv = (uint32_t)V;
            if (dt == PB200_U8) mode_keys_kernel<uint8_t><<<blocks, 256, 0, st>>>(sts, nv, si2, n, ra.src, sstride, ra.src_aligned, mk);
            else if (dt == PB200_I8) mode_keys_kernel<int8_t><<<blocks, 256, 0, st>>>(sts, nv, si2, n, ra.src, sstride, ra.src_aligned, mk);
            else if (dt == PB200_I16) mode_keys_kernel<int16_t><<<blocks, 256, 0, st>>>(sts, nv, si2, n, ra.src, sstride, ra.src_aligned, mk);
            else mode_keys_kernel<uint16_t><<<blocks, 256, 0, st>>>(sts, nv, si2, n, ra.src, sstride, ra.src_aligned, mk);
            bool in_alt = false;
            int src_rc = radix_sort_u64(ctx, mk, (unsigned long long*)d_mode_keys2.p, nullptr, nullptr, n, 0, 16 + (int)bits_for(V + 1), &in_alt);
            if (src_rc < 0) return fail(src_rc);
            const unsigned long long* sorted_mk = in_alt ? (const unsigned long long*)d_mode_keys2.p : mk;
            cudaMemsetAsync(d_best.p, 0, (V + 1) * 8, st);
            run_vote_kernel<<<blocks, 256, 0, st>>>(sorted_mk, n, (unsigned long long*)d_best.p);
            mode_decode_kernel<<<blocks, 256, 0, st>>>((const unsigned long long*)d_best.p, V, ra.dst, ra.dst_size, rules[a]->kind == R_MODE_BOOL ? 1 : 0,
                                                           (dt == PB200_I8 || dt == PB200_I16) ? 32768ll : 0ll);
            g_launches += 3;
        } else {
            PB_PHASE(ctx, "voxel.reduce");
            launch_reduce(*rules[a], ra, st);
        }
    }
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(cuda_error(e, "voxel reduce"));
    }
    // packed voxel keys stay on the device; pb200_result_buffer_voxel_keys unpacks them on demand
    res->d_packed_keys = d_vkeys.p;
    d_vkeys.p = nullptr;  // ownership moves to the result
    res->bits_y = bits_y;
    res->bits_z = bits_z;
    // ---- hand the result over in the requested memory layout / space -------------------------------------------
    pb200_buffer_desc stage;
    stage.layout = dst_layout;
    stage.kind = PB200_COLUMNAR;
    stage.memspace = PB200_DEVICE;
    stage.len = V;
    stage.aos = nullptr;
    stage.columns = d_out.data();
    if (dst_kind == PB200_COLUMNAR && dst_memspace == PB200_DEVICE) {
        // device-resident columnar result: nothing to wait for -- the temporaries are stream-ordered (DevTmp) and the
        // result is valid for work queued on the context's stream
        res->columns = d_out;
        *out = res;
        return PB200_OK;
    }
    cudaStreamSynchronize(st);
    // allocate the final storage
    if (dst_kind == PB200_INTERLEAVED) {
        const size_t bytes = (size_t)(V * dst_layout->size) + 16;
        if (dst_memspace == PB200_DEVICE) { if (cache_alloc(ctx, &res->aos, bytes) != cudaSuccess) return fail(set_error(PB200_ERR_OOM, "out of device memory")); cudaMemsetAsync(res->aos, 0, bytes, st); }
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
    for (void* p : d_out) if (p) cache_free(ctx, p);
    if (rc < 0) {
        std::fill(d_out.begin(), d_out.end(), nullptr);
        pb200_result_buffer_destroy(res);
        return rc;
    }
    *out = res;
    return PB200_OK;
}

// positions of a buffer as a device (base, stride) view; host buffers are staged into `staged`
static int voxel_positions(pb200_ctx* ctx, const pb200_buffer_desc* buf, DevTmp* staged, const uint8_t** base, uint64_t* stride) {
    const int pi = pb200_layout_index_of(buf->layout, "Position3D", PB200_VEC3F64);
    if (pi < 0)
        return set_error(PB200_ERR_ATTR_NOT_FOUND, "The PointBuffer does not have the attribute attributes::POSITION_3D which is needed for the creation of the voxel grid.");
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
    return PB200_OK;
}

int pb200_voxelgrid_partials(pb200_ctx* ctx, const pb200_buffer_desc* src, double lx, double ly, double lz,
                             const double global_min[3], const double global_max[3], pb200_voxel_partials** out) {
    if (!ctx || !out || !global_min || !global_max) return set_error(PB200_ERR_INVALID, "null argument");
    *out = nullptr;
    PB_TRY(validate_desc(src, "source buffer"));
    PB_DEVICE(ctx);
    const double leaf[3] = {lx, ly, lz};
    for (int c = 0; c < 3; ++c) {
        if (!(leaf[c] > 0.0)) return set_error(PB200_ERR_INVALID, "leaf sizes must be positive (the reference would not terminate)");
        if (!(global_min[c] <= global_max[c])) return set_error(PB200_ERR_INVALID, "global bounds: min > max (AABB::from_min_max panics)");
    }
    if (src->len > 0xFFFFFFFFull) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^32-1 points per call");
    cudaStream_t st = ctx->stream;
    pb200_voxel_partials* res = new pb200_voxel_partials();
    res->ctx = ctx;
    auto fail = [&](int rc) { pb200_voxel_partials_destroy(res); return rc; };
    if (src->len == 0) {  // an empty shard contributes nothing; grid geometry is still reported
        for (int c = 0; c < 3; ++c) {
            uint64_t cnt = 0;
            for (double cur = global_min[c]; cur < global_max[c]; cur += leaf[c])
                if (++cnt > (1u << 21)) return fail(set_error(PB200_ERR_UNSUPPORTED, "more than 2^21 voxels along one axis"));
            res->cells[c] = cnt;
        }
        res->bits_x = bits_for(res->cells[0]); res->bits_y = bits_for(res->cells[1]); res->bits_z = bits_for(res->cells[2]);
        *out = res;
        return PB200_OK;
    }
    DevTmp staged;
    const uint8_t* ppos = nullptr;
    uint64_t pstride = 0;
    int rc = voxel_positions(ctx, src, &staged, &ppos, &pstride);
    if (rc < 0) return fail(rc);
    VoxelIndex vi;
    rc = voxel_index_sort(ctx, ppos, pstride, src->len, global_min, global_max, leaf, &vi);
    if (rc < 0) return fail(rc);
    res->len = vi.V;
    res->bits_x = vi.bits_x; res->bits_y = vi.bits_y; res->bits_z = vi.bits_z;
    for (int c = 0; c < 3; ++c) res->cells[c] = vi.cells[c];
    if (cache_alloc(ctx, &res->counts, vi.V * 4 + 4) != cudaSuccess || cache_alloc(ctx, &res->sums, vi.V * 24 + 8) != cudaSuccess)
        return fail(set_error(PB200_ERR_OOM, "out of device memory"));
    rc = voxel_index_emit(ctx, &vi, 1, (double*)res->sums, (uint32_t*)res->counts, false);  // keys + per-voxel sums and counts
    if (rc < 0) return fail(rc);
    res->keys = vi.voxel_keys.p;
    vi.voxel_keys.p = nullptr;  // ownership moves to the result
    if (src->memspace == PB200_HOST) cudaStreamSynchronize(st);
    *out = res;
    return PB200_OK;
}

int pb200_voxelgrid_merge_partials(pb200_ctx* ctx, const uint64_t* keys, const uint32_t* counts, const double* sums, uint64_t m,
                                   uint32_t bits_x, uint32_t bits_y, uint32_t bits_z, pb200_voxel_partials** out) {
    if (!ctx || !out) return set_error(PB200_ERR_INVALID, "null argument");
    *out = nullptr;
    if (m && (!keys || !counts || !sums)) return set_error(PB200_ERR_INVALID, "null partial arrays");
    if (m > 0xFFFFFFFEull) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^32-2 partial voxels per call");
    if (bits_x + bits_y + bits_z > 64 || bits_x + bits_y + bits_z == 0) return set_error(PB200_ERR_INVALID, "bad key widths");
    PB_DEVICE(ctx);
    cudaStream_t st = ctx->stream;
    pb200_voxel_partials* res = new pb200_voxel_partials();
    res->ctx = ctx;
    res->bits_x = bits_x; res->bits_y = bits_y; res->bits_z = bits_z;
    if (m == 0) { *out = res; return PB200_OK; }
    auto fail = [&](int rc) { pb200_voxel_partials_destroy(res); return rc; };
    DevTmp d_keys1, d_keys2, d_idx, d_idx2, d_tiles, d_starts;
    int rc = PB200_OK;
    auto run = [&]() -> int {
        PB_CUDA(d_keys1.alloc(st, m * 8)); PB_CUDA(d_keys2.alloc(st, m * 8)); PB_CUDA(d_idx.alloc(st, m * 4)); PB_CUDA(d_idx2.alloc(st, m * 4));
        const unsigned long long cap = (unsigned long long)ctx->sm_count * 16;
        const unsigned blocks = (unsigned)(((m + 255) / 256) < cap ? ((m + 255) / 256) : cap);
        iota_kernel<<<blocks, 256, 0, st>>>((uint32_t*)d_idx.p, m);
        PB_CUDA(cudaMemcpyAsync(d_keys1.p, keys, m * 8, cudaMemcpyDeviceToDevice, st));  // the caller's keys stay intact
        const int end_bit = (int)(bits_x + bits_y + bits_z);
        bool in_alt = false;
        PB_TRY(radix_sort_u64(ctx, (unsigned long long*)d_keys1.p, (unsigned long long*)d_keys2.p, (uint32_t*)d_idx.p, (uint32_t*)d_idx2.p,
                              m, 0, end_bit, &in_alt));
        if (!in_alt) { std::swap(d_keys1.p, d_keys2.p); std::swap(d_idx.p, d_idx2.p); }  // sorted data in d_keys2 / d_idx2
        const uint32_t n_tiles = (uint32_t)((m + HT_TILE - 1) / HT_TILE);
        PB_CUDA(d_tiles.alloc(st, ((size_t)n_tiles + 1) * 4));
        uint32_t* d_total = (uint32_t*)d_tiles.p + n_tiles;
        heads_count_kernel<<<n_tiles, HT_THREADS, 0, st>>>((const unsigned long long*)d_keys2.p, m, 0u, (uint32_t*)d_tiles.p);
        PB_TRY(exclusive_scan_u32(ctx, (uint32_t*)d_tiles.p, n_tiles, d_total));
        uint32_t* h_total = (uint32_t*)ctx->h_scratch;
        PB_CUDA(cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        const uint64_t V = *h_total;
        res->len = V;
        PB_CUDA(d_starts.alloc(st, (V + 1) * 4));
        PB_CUDA(cache_alloc(ctx, &res->keys, V * 8 + 8));
        PB_CUDA(cache_alloc(ctx, &res->counts, V * 4 + 4));
        PB_CUDA(cache_alloc(ctx, &res->sums, V * 24 + 8));
        heads_emit_kernel<<<n_tiles, HT_THREADS, 0, st>>>((const unsigned long long*)d_keys2.p, m, 0u, (const uint32_t*)d_tiles.p, (uint32_t)V,
                                                          (uint32_t*)d_starts.p, (unsigned long long*)res->keys, nullptr);
        partials_merge_kernel<<<(unsigned)((V + 127) / 128), 128, 0, st>>>((const uint32_t*)d_starts.p, (const uint32_t*)d_idx2.p, V, counts, sums,
                                                                           (uint32_t*)res->counts, (double*)res->sums);
        g_launches += 5;
        PB_CUDA(cudaGetLastError());
        return PB200_OK;
    };
    rc = run();
    if (rc < 0) return fail(rc);
    *out = res;
    return PB200_OK;
}

// ---- sharded grid with attributes -----------------------------------------------------------------------------
namespace {

// how the attributes of a filtered layout travel between shards: Position3D through (count, sums); mean and max-pool
// attributes as columns; "most common value" attributes as run lists.  The same walk over dst_layout on both sides
// (partials, merge) yields the same column / list numbering.
struct AttrPlan {
    std::vector<const Rule*> rules;
    std::vector<int> col0;      // first column of attribute a, -1 if it has none
    std::vector<int> mode_no;   // list number of attribute a, -1 if it is not a mode attribute
    uint32_t n_col = 0, n_modes = 0;
    ColOps ops;
};

int plan_attrs(const pb200_layout* dst_layout, AttrPlan* pl) {
    static const char* WAVE[] = {"WaveformDataOffset", "WaveformPacketSize", "WaveformParameters", "WavePacketDescriptorIndex", "ReturnPointWaveformLocation"};
    static const uint32_t WAVE_T[] = {PB200_U64, PB200_U32, PB200_VEC3F32, PB200_U8, PB200_F32};
    for (int w = 0; w < 5; ++w)
        if (pb200_layout_index_of(dst_layout, WAVE[w], WAVE_T[w]) >= 0) return set_error(PB200_ERR_UNSUPPORTED, "Waveform data currently not supported!");
    memset(&pl->ops, 0, sizeof pl->ops);
    for (const auto& a : dst_layout->attrs) {
        const Rule* r = find_rule(a);
        if (!r) return set_error(PB200_ERR_UNSUPPORTED, "attribute is non-standard which is not supported currently: %s", a.name);
        pl->rules.push_back(r);
        int c0 = -1, mn = -1;
        uint32_t w = 0;
        bool mx = false;
        switch (r->kind) {
            case R_MEAN_VEC_F64: break;  // position: (count, sums)
            case R_MEAN_U16: w = 1; break;
            case R_MEAN_VEC_U16: case R_MEAN_VEC_F32: w = 3; break;
            case R_MAX_U8: case R_MAX_F64: case R_MAX_U64: w = 1; mx = true; break;
            case R_MODE: case R_MODE_BOOL: mn = (int)pl->n_modes++; break;
        }
        if (w) {
            if (pl->n_col + w > 64) return set_error(PB200_ERR_UNSUPPORTED, "more than 64 partial columns");
            c0 = (int)pl->n_col;
            for (uint32_t k = 0; k < w; ++k) pl->ops.is_max[pl->n_col + k] = mx ? 1 : 0;
            pl->n_col += w;
        }
        pl->col0.push_back(c0);
        pl->mode_no.push_back(mn);
    }
    return PB200_OK;
}

// run starts / run keys of an ascending key array (shift 0): the heads machinery of the voxel boundaries
int key_runs(pb200_ctx* ctx, const unsigned long long* sorted_keys, uint64_t m, DevTmp* starts, DevTmp* run_keys, uint64_t* n_runs) {
    cudaStream_t st = ctx->stream;
    DevTmp d_tiles;
    const uint32_t n_tiles = (uint32_t)((m + HT_TILE - 1) / HT_TILE);
    PB_CUDA(d_tiles.alloc(st, ((size_t)n_tiles + 1) * 4));
    uint32_t* d_total = (uint32_t*)d_tiles.p + n_tiles;
    heads_count_kernel<<<n_tiles, HT_THREADS, 0, st>>>(sorted_keys, m, 0u, (uint32_t*)d_tiles.p);
    PB_TRY(exclusive_scan_u32(ctx, (uint32_t*)d_tiles.p, n_tiles, d_total));
    uint32_t* h_total = (uint32_t*)ctx->h_scratch;
    PB_CUDA(cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    const uint64_t R = *h_total;
    *n_runs = R;
    PB_CUDA(starts->alloc(st, (R + 1) * 4));
    PB_CUDA(run_keys->alloc(st, R * 8 + 8));
    heads_emit_kernel<<<n_tiles, HT_THREADS, 0, st>>>(sorted_keys, m, 0u, (const uint32_t*)d_tiles.p, (uint32_t)R, (uint32_t*)starts->p,
                                                      (unsigned long long*)run_keys->p, nullptr);
    g_launches += 3;
    PB_CUDA(cudaGetLastError());
    return PB200_OK;
}

unsigned grid_cap(pb200_ctx* ctx, uint64_t n) {
    const unsigned long long cap = (unsigned long long)ctx->sm_count * 16, want = (n + 255) / 256;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

int pb200_voxelgrid_partials_layout(pb200_ctx* ctx, const pb200_buffer_desc* src, double lx, double ly, double lz,
                                    const double global_min[3], const double global_max[3], const pb200_layout* dst_layout,
                                    pb200_voxel_partials** out) {
    if (!ctx || !out || !global_min || !global_max || !dst_layout) return set_error(PB200_ERR_INVALID, "null argument");
    *out = nullptr;
    PB_TRY(validate_desc(src, "source buffer"));
    PB_DEVICE(ctx);
    if (src->memspace != PB200_DEVICE) return set_error(PB200_ERR_UNSUPPORTED, "attribute partials need a device-resident shard");
    const double leaf[3] = {lx, ly, lz};
    for (int c = 0; c < 3; ++c) {
        if (!(leaf[c] > 0.0)) return set_error(PB200_ERR_INVALID, "leaf sizes must be positive (the reference would not terminate)");
        if (!(global_min[c] <= global_max[c])) return set_error(PB200_ERR_INVALID, "global bounds: min > max (AABB::from_min_max panics)");
    }
    if (src->len > 0xFFFFFFFFull) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^32-1 points per call");
    AttrPlan ap;
    PB_TRY(plan_attrs(dst_layout, &ap));
    std::vector<int> src_idx;
    for (size_t a = 0; a < ap.rules.size(); ++a) {
        const int si = pb200_layout_index_of(src->layout, ap.rules[a]->name, ap.rules[a]->dtype);
        if (si < 0) return set_error(PB200_ERR_ATTR_NOT_FOUND, "source buffer has no attribute %s with the default datatype", ap.rules[a]->name);
        src_idx.push_back(si);
    }
    cudaStream_t st = ctx->stream;
    pb200_voxel_partials* res = new pb200_voxel_partials();
    res->ctx = ctx;
    res->n_col = ap.n_col;
    res->n_modes = ap.n_modes;
    memcpy(res->col_is_max, ap.ops.is_max, 64);
    auto fail = [&](int rc) { pb200_voxel_partials_destroy(res); return rc; };
    if (src->len == 0) {  // an empty shard contributes nothing; grid geometry is still reported
        for (int c = 0; c < 3; ++c) {
            uint64_t cnt = 0;
            for (double cur = global_min[c]; cur < global_max[c]; cur += leaf[c])
                if (++cnt > (1u << 21)) return fail(set_error(PB200_ERR_UNSUPPORTED, "more than 2^21 voxels along one axis"));
            res->cells[c] = cnt;
        }
        res->bits_x = bits_for(res->cells[0]); res->bits_y = bits_for(res->cells[1]); res->bits_z = bits_for(res->cells[2]);
        *out = res;
        return PB200_OK;
    }
    DevTmp staged;
    const uint8_t* ppos = nullptr;
    uint64_t pstride = 0;
    int rc = voxel_positions(ctx, src, &staged, &ppos, &pstride);
    if (rc < 0) return fail(rc);
    VoxelIndex vi;
    rc = voxel_index_sort(ctx, ppos, pstride, src->len, global_min, global_max, leaf, &vi);
    if (rc < 0) return fail(rc);
    const uint64_t V = vi.V, n = src->len;
    res->len = V;
    res->bits_x = vi.bits_x; res->bits_y = vi.bits_y; res->bits_z = vi.bits_z;
    for (int c = 0; c < 3; ++c) res->cells[c] = vi.cells[c];
    if (vi.bits_x + vi.bits_y + vi.bits_z > 48 && ap.n_modes) return fail(set_error(PB200_ERR_UNSUPPORTED, "mode attributes need voxel keys of at most 48 bits"));
    if (cache_alloc(ctx, &res->counts, V * 4 + 4) != cudaSuccess || cache_alloc(ctx, &res->sums, V * 24 + 8) != cudaSuccess)
        return fail(set_error(PB200_ERR_OOM, "out of device memory"));
    const bool need_index = ap.n_col + ap.n_modes > 0;
    rc = voxel_index_emit(ctx, &vi, 1, (double*)res->sums, (uint32_t*)res->counts, need_index);
    if (rc < 0) return fail(rc);
    auto attr_ptr = [&](int idx, uint64_t* stride) -> const uint8_t* {
        const pb200_attr& a = src->layout->attrs[(size_t)idx];
        if (src->kind == PB200_INTERLEAVED) { *stride = src->layout->size; return (const uint8_t*)src->aos + a.offset; }
        *stride = a.size;
        return (const uint8_t*)src->columns[idx];
    };
    if (ap.n_col && cache_alloc(ctx, &res->cols, (size_t)V * ap.n_col * 8 + 8) != cudaSuccess) return fail(set_error(PB200_ERR_OOM, "out of device memory"));
    DevTmp d_mk, d_mk2;
    if (ap.n_modes && (d_mk.alloc(st, n * 8) != cudaSuccess || d_mk2.alloc(st, n * 8) != cudaSuccess)) return fail(set_error(PB200_ERR_OOM, "out of device memory"));
    for (size_t a = 0; a < ap.rules.size(); ++a) {
        const Rule& r = *ap.rules[a];
        uint64_t sstride = 0;
        const uint8_t* sp = attr_ptr(src_idx[a], &sstride);
        const uint64_t comp = pb200_dtype_size(is_cast_vec3(r.dtype) ? vec3_component(r.dtype) : r.dtype, 0);
        const int aligned = (((uintptr_t)sp % comp) == 0 && (sstride % comp) == 0) ? 1 : 0;
        if (ap.col0[a] >= 0) {
            PB_PHASE(ctx, "voxel.partial_columns");
            ReduceArgs ra;
            ra.starts = (const uint32_t*)vi.starts.p;
            ra.sorted_idx = (const uint32_t*)vi.sorted_idx.p;
            ra.n_voxels = V;
            ra.src = sp;
            ra.src_stride = sstride;
            ra.dst = nullptr;
            ra.dst_size = 0;
            ra.src_aligned = aligned;
            ra.partial = (double*)res->cols + ap.col0[a];
            ra.partial_stride = ap.n_col;
            launch_reduce(r, ra, st);
        } else if (ap.mode_no[a] >= 0) {
            PB_PHASE(ctx, "voxel.partial_mode_runs");
            const int mno = ap.mode_no[a];
            unsigned long long* mk = (unsigned long long*)d_mk.p;
            const uint32_t *sts = (const uint32_t*)vi.starts.p, *si2 = (const uint32_t*)vi.sorted_idx.p;
            const unsigned blocks = grid_cap(ctx, n);
            if (r.dtype == PB200_U8) mode_keys_kernel<uint8_t><<<blocks, 256, 0, st>>>(sts, (uint32_t)V, si2, n, sp, sstride, aligned, mk);
            else if (r.dtype == PB200_I8) mode_keys_kernel<int8_t><<<blocks, 256, 0, st>>>(sts, (uint32_t)V, si2, n, sp, sstride, aligned, mk);
            else if (r.dtype == PB200_I16) mode_keys_kernel<int16_t><<<blocks, 256, 0, st>>>(sts, (uint32_t)V, si2, n, sp, sstride, aligned, mk);
            else mode_keys_kernel<uint16_t><<<blocks, 256, 0, st>>>(sts, (uint32_t)V, si2, n, sp, sstride, aligned, mk);
            g_launches++;
            bool in_alt = false;
            rc = radix_sort_u64(ctx, mk, (unsigned long long*)d_mk2.p, nullptr, nullptr, n, 0, 16 + (int)bits_for(V + 1), &in_alt);
            if (rc < 0) return fail(rc);
            const unsigned long long* sorted_mk = in_alt ? (const unsigned long long*)d_mk2.p : mk;
            DevTmp run_starts, run_keys;
            uint64_t R = 0;
            rc = key_runs(ctx, sorted_mk, n, &run_starts, &run_keys, &R);
            if (rc < 0) return fail(rc);
            if (cache_alloc(ctx, &res->mode_keys[mno], R * 8 + 8) != cudaSuccess || cache_alloc(ctx, &res->mode_counts[mno], R * 4 + 4) != cudaSuccess)
                return fail(set_error(PB200_ERR_OOM, "out of device memory"));
            res->mode_len[mno] = R;
            mode_runs_kernel<<<grid_cap(ctx, R), 256, 0, st>>>((const unsigned long long*)run_keys.p, (const uint32_t*)run_starts.p, R,
                                                              (const unsigned long long*)vi.voxel_keys.p, (unsigned long long*)res->mode_keys[mno],
                                                              (uint32_t*)res->mode_counts[mno]);
            g_launches++;
        }
    }
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(cuda_error(e, "attribute partials"));
    }
    res->keys = vi.voxel_keys.p;
    vi.voxel_keys.p = nullptr;  // ownership moves to the result
    *out = res;
    return PB200_OK;
}

int pb200_voxel_partials_get_attrs(const pb200_voxel_partials* p, pb200_voxel_attr_partials_desc* out) {
    if (!p || !out) return set_error(PB200_ERR_INVALID, "null argument");
    memset(out, 0, sizeof(*out));
    out->n_columns = p->n_col;
    out->n_modes = p->n_modes;
    out->columns = (const double*)p->cols;
    memcpy(out->column_is_max, p->col_is_max, 64);
    for (uint32_t a = 0; a < PB200_MAX_ATTRIBUTES; ++a) {
        out->mode_len[a] = p->mode_len[a];
        out->mode_keys[a] = (const uint64_t*)p->mode_keys[a];
        out->mode_counts[a] = (const uint32_t*)p->mode_counts[a];
    }
    return PB200_OK;
}

int pb200_voxelgrid_merge_partials_layout(pb200_ctx* ctx, const pb200_layout* dst_layout, const pb200_voxel_partials_desc* pos,
                                          const pb200_voxel_attr_partials_desc* attrs, pb200_result_buffer** out) {
    if (!ctx || !dst_layout || !pos || !attrs || !out) return set_error(PB200_ERR_INVALID, "null argument");
    *out = nullptr;
    const uint64_t m = pos->len;
    if (m && (!pos->keys || !pos->counts || !pos->sums)) return set_error(PB200_ERR_INVALID, "null partial arrays");
    if (m > 0xFFFFFFFEull) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^32-2 partial voxels per call");
    const int key_bits = (int)(pos->bits_x + pos->bits_y + pos->bits_z);
    if (key_bits > 64 || key_bits == 0) return set_error(PB200_ERR_INVALID, "bad key widths");
    AttrPlan ap;
    PB_TRY(plan_attrs(dst_layout, &ap));
    if (ap.n_col != attrs->n_columns || ap.n_modes != attrs->n_modes) return set_error(PB200_ERR_LAYOUT_MISMATCH, "attribute partials do not belong to this layout");
    if (ap.n_col && m && !attrs->columns) return set_error(PB200_ERR_INVALID, "null partial columns");
    if (ap.n_modes && key_bits > 48) return set_error(PB200_ERR_UNSUPPORTED, "mode attributes need voxel keys of at most 48 bits");
    PB_DEVICE(ctx);
    cudaStream_t st = ctx->stream;
    pb200_result_buffer* res = new pb200_result_buffer();
    res->ctx = ctx;
    res->layout = *dst_layout;
    res->kind = PB200_COLUMNAR;
    res->memspace = PB200_DEVICE;
    res->bits_y = pos->bits_y;
    res->bits_z = pos->bits_z;
    res->columns.assign(dst_layout->attrs.size(), nullptr);
    auto fail = [&](int rc) { pb200_result_buffer_destroy(res); return rc; };
    if (m == 0) { *out = res; return PB200_OK; }
    // ---- voxels: sort the concatenated partial keys (stable: source-rank order inside a voxel), find the runs -------------
    DevTmp d_keys1, d_keys2, d_idx, d_idx2, d_starts, d_vkeys, d_counts, d_sums, d_cols;
    uint64_t V = 0;
    auto run = [&]() -> int {
        PB_PHASE(ctx, "voxel.merge");
        PB_CUDA(d_keys1.alloc(st, m * 8)); PB_CUDA(d_keys2.alloc(st, m * 8)); PB_CUDA(d_idx.alloc(st, m * 4)); PB_CUDA(d_idx2.alloc(st, m * 4));
        iota_kernel<<<grid_cap(ctx, m), 256, 0, st>>>((uint32_t*)d_idx.p, m);
        PB_CUDA(cudaMemcpyAsync(d_keys1.p, pos->keys, m * 8, cudaMemcpyDeviceToDevice, st));
        bool in_alt = false;
        PB_TRY(radix_sort_u64(ctx, (unsigned long long*)d_keys1.p, (unsigned long long*)d_keys2.p, (uint32_t*)d_idx.p, (uint32_t*)d_idx2.p,
                              m, 0, key_bits, &in_alt));
        if (!in_alt) { std::swap(d_keys1.p, d_keys2.p); std::swap(d_idx.p, d_idx2.p); }
        PB_TRY(key_runs(ctx, (const unsigned long long*)d_keys2.p, m, &d_starts, &d_vkeys, &V));
        PB_CUDA(d_counts.alloc(st, V * 4 + 4)); PB_CUDA(d_sums.alloc(st, V * 24 + 8));
        partials_merge_kernel<<<(unsigned)((V + 127) / 128), 128, 0, st>>>((const uint32_t*)d_starts.p, (const uint32_t*)d_idx2.p, V, pos->counts,
                                                                           pos->sums, (uint32_t*)d_counts.p, (double*)d_sums.p);
        if (ap.n_col) {
            PB_CUDA(d_cols.alloc(st, (size_t)V * ap.n_col * 8 + 8));
            partials_merge_columns_kernel<<<(unsigned)((V + 127) / 128), 128, 0, st>>>((const uint32_t*)d_starts.p, (const uint32_t*)d_idx2.p, V,
                                                                                       attrs->columns, ap.n_col, ap.ops, (double*)d_cols.p);
        }
        g_launches += 3;
        PB_CUDA(cudaGetLastError());
        return PB200_OK;
    };
    int rc = run();
    if (rc < 0) return fail(rc);
    res->len = V;
    for (size_t a = 0; a < dst_layout->attrs.size(); ++a)
        if (cache_alloc(ctx, &res->columns[a], (size_t)(V * dst_layout->attrs[a].size) + 16) != cudaSuccess) return fail(set_error(PB200_ERR_OOM, "out of device memory"));
    // ---- attribute values ------------------------------------------------------------------------------------------------
    DevTmp d_best;
    if (ap.n_modes && d_best.alloc(st, (V + 1) * 8) != cudaSuccess) return fail(set_error(PB200_ERR_OOM, "out of device memory"));
    for (size_t a = 0; a < ap.rules.size(); ++a) {
        const Rule& r = *ap.rules[a];
        uint8_t* dst = (uint8_t*)res->columns[a];
        const uint32_t dsz = (uint32_t)dst_layout->attrs[a].size;
        if (r.kind == R_MEAN_VEC_F64) {
            partials_centroid_kernel<<<grid_cap(ctx, V), 256, 0, st>>>((const uint32_t*)d_counts.p, (const double*)d_sums.p, V, (double*)dst);
            g_launches++;
        } else if (ap.col0[a] >= 0) {
            finalize_columns_kernel<<<grid_cap(ctx, V), 256, 0, st>>>((int)r.kind, (const double*)d_cols.p, ap.n_col, (uint32_t)ap.col0[a],
                                                                      (const uint32_t*)d_counts.p, V, dst, dsz);
            g_launches++;
        } else {
            PB_PHASE(ctx, "voxel.merge_modes");
            const int mno = ap.mode_no[a];
            const uint64_t rl = attrs->mode_len[mno];
            cudaMemsetAsync(d_best.p, 0, (V + 1) * 8, st);
            if (rl) {
                if (!attrs->mode_keys[mno] || !attrs->mode_counts[mno]) return fail(set_error(PB200_ERR_INVALID, "null run list"));
                if (rl > 0xFFFFFFFEull) return fail(set_error(PB200_ERR_UNSUPPORTED, "more than 2^32-2 runs per call"));
                DevTmp k1, k2, i1, i2, rs, rk;
                if (k1.alloc(st, rl * 8) != cudaSuccess || k2.alloc(st, rl * 8) != cudaSuccess || i1.alloc(st, rl * 4) != cudaSuccess ||
                    i2.alloc(st, rl * 4) != cudaSuccess)
                    return fail(set_error(PB200_ERR_OOM, "out of device memory"));
                iota_kernel<<<grid_cap(ctx, rl), 256, 0, st>>>((uint32_t*)i1.p, rl);
                cudaMemcpyAsync(k1.p, attrs->mode_keys[mno], rl * 8, cudaMemcpyDeviceToDevice, st);
                bool in_alt = false;
                rc = radix_sort_u64(ctx, (unsigned long long*)k1.p, (unsigned long long*)k2.p, (uint32_t*)i1.p, (uint32_t*)i2.p, rl, 0, key_bits + 16, &in_alt);
                if (rc < 0) return fail(rc);
                if (!in_alt) { std::swap(k1.p, k2.p); std::swap(i1.p, i2.p); }
                uint64_t R = 0;
                rc = key_runs(ctx, (const unsigned long long*)k2.p, rl, &rs, &rk, &R);
                if (rc < 0) return fail(rc);
                mode_merge_vote_kernel<<<grid_cap(ctx, R), 256, 0, st>>>((const uint32_t*)rs.p, (const uint32_t*)i2.p, (const unsigned long long*)rk.p, R,
                                                                        attrs->mode_counts[mno], (const unsigned long long*)d_vkeys.p, V,
                                                                        (unsigned long long*)d_best.p);
                g_launches += 2;
            }
            mode_decode_kernel<<<grid_cap(ctx, V), 256, 0, st>>>((const unsigned long long*)d_best.p, V, dst, dsz, r.kind == R_MODE_BOOL ? 1 : 0,
                                                                (r.dtype == PB200_I8 || r.dtype == PB200_I16) ? 32768ll : 0ll);
            g_launches++;
        }
    }
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(cuda_error(e, "merge of attribute partials"));
    }
    res->d_packed_keys = d_vkeys.p;
    d_vkeys.p = nullptr;  // ownership moves to the result
    *out = res;
    return PB200_OK;
}

int pb200_voxel_partials_get(const pb200_voxel_partials* p, pb200_voxel_partials_desc* out) {
    if (!p || !out) return set_error(PB200_ERR_INVALID, "null argument");
    out->len = p->len;
    out->keys = (const uint64_t*)p->keys;
    out->counts = (const uint32_t*)p->counts;
    out->sums = (const double*)p->sums;
    out->bits_x = p->bits_x; out->bits_y = p->bits_y; out->bits_z = p->bits_z;
    for (int c = 0; c < 3; ++c) out->cells[c] = p->cells[c];
    return PB200_OK;
}

int pb200_voxel_partials_centroids(const pb200_voxel_partials* p, double* positions_out) {
    if (!p || (!positions_out && p->len)) return set_error(PB200_ERR_INVALID, "null argument");
    if (p->len == 0) return PB200_OK;
    PB_DEVICE(p->ctx);
    const unsigned long long cap = (unsigned long long)p->ctx->sm_count * 16;
    const unsigned blocks = (unsigned)(((p->len + 255) / 256) < cap ? ((p->len + 255) / 256) : cap);
    partials_centroid_kernel<<<blocks, 256, 0, p->ctx->stream>>>((const uint32_t*)p->counts, (const double*)p->sums, p->len, positions_out);
    g_launches++;
    PB_CUDA(cudaGetLastError());
    return PB200_OK;
}

void pb200_voxel_partials_destroy(pb200_voxel_partials* p) {
    if (!p) return;
    if (p->ctx) cudaSetDevice(p->ctx->device);
    if (p->ctx) {
        cache_free(p->ctx, p->keys); cache_free(p->ctx, p->counts); cache_free(p->ctx, p->sums); cache_free(p->ctx, p->cols);
        for (uint32_t a = 0; a < PB200_MAX_ATTRIBUTES; ++a) { cache_free(p->ctx, p->mode_keys[a]); cache_free(p->ctx, p->mode_counts[a]); }
    }
    delete p;
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
    PB_DEVICE(r->ctx);
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
    if (r->d_packed_keys && r->ctx) cache_free(r->ctx, r->d_packed_keys);
    if (r->memspace == PB200_DEVICE) {
        if (r->ctx) { cache_free(r->ctx, r->aos); for (void* p : r->columns) cache_free(r->ctx, p); }
    } else {
        free(r->aos);
        for (void* p : r->columns) free(p);
    }
    delete r;
}

}  // extern "C"
