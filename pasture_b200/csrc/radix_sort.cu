// radix_sort.cu -- stable LSD radix sort of 64-bit keys (optionally with a 32-bit payload), 8-bit digits, ONE sweep
// over the data per digit: the sort behind the voxel grid (K8: packed voxel key | point index, voxel_grid.rs:141-152
// keeps its voxels sorted) and the LBVH (K8: Morton code, point index).
//
//   1. radix_histogram_kernel   one read of the keys -> digit histograms of ALL passes (shared-memory atomics, then
//                               one global atomic per bin and CTA); a tiny kernel turns them into bin bases
//   2. radix_onesweep_kernel    per pass.  A CTA takes the next tile (ticket counter: tiles are claimed in order, so a
//                               CTA only ever waits for tiles that are already running), ranks its 8192 keys by digit
//                               (per-bit ballots + warp-private digit counters, stable), publishes the tile's digit
//                               counts, resolves its global offsets by DECOUPLED LOOK-BACK over the preceding tiles
//                               (one thread per digit; flag and value share one 32-bit word, so a single relaxed load
//                               is a consistent snapshot), reorders the tile in shared memory and writes each digit's
//                               run to its final place: every digit run is a contiguous burst.
// Algorithmic bytes per key and pass: 8 read + 8 written (+ 4 + 4 with a payload); the histogram adds one 8-byte read.
#include "internal.h"

namespace pb200 {

namespace rs {

// Digits are 8 or 9 bits wide: a 9-bit pass (512 bins, one look-back thread per bin = the whole CTA) costs about the same
// as an 8-bit one, so whenever ceil(bits / 9) < ceil(bits / 8) the plan uses 9-bit digits and saves a whole sweep over the
// data (the 34 voxel-key bits of C3: 9+9+8+8 = 4 passes instead of 5; the 63 Morton bits: 7 instead of 8).
constexpr int MAX_RADIX_BITS = 9, MAX_RADIX = 1 << MAX_RADIX_BITS;
constexpr int THREADS = 512, WARPS = THREADS / 32, IPT = 16, TILE = THREADS * IPT;  // 8192 keys per tile
constexpr int MAX_PASSES = 8;
constexpr int LOOK = 8;  // predecessors fetched per look-back batch
constexpr uint32_t FLAG_LOCAL = 1u << 30, FLAG_INCLUSIVE = 2u << 30, VALUE_MASK = (1u << 30) - 1u;

struct Passes {
    int n_passes;
    int shift[MAX_PASSES];
    uint32_t mask[MAX_PASSES];
};

__global__ void __launch_bounds__(512) radix_histogram_kernel(const unsigned long long* __restrict__ keys, unsigned long long n,
                                                              Passes ps, uint32_t* __restrict__ hist /* [passes][256] */) {
    __shared__ uint32_t s_hist[MAX_PASSES * MAX_RADIX];
    constexpr int RADIX = MAX_RADIX;  // histogram rows are MAX_RADIX wide whatever the digit width of a pass
    for (int i = threadIdx.x; i < ps.n_passes * RADIX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    auto count = [&](unsigned long long k) {
#pragma unroll
        for (int p = 0; p < MAX_PASSES; ++p)
            if (p < ps.n_passes) atomicAdd(&s_hist[p * RADIX + (uint32_t)((k >> ps.shift[p]) & ps.mask[p])], 1u);
    };
    // two keys per 16-byte load (cudaMalloc'd key arrays are 16-byte aligned; an odd tail key is counted separately)
    const unsigned long long pairs = ((reinterpret_cast<uintptr_t>(keys) & 15) == 0) ? n / 2 : 0;
    const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(keys);
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += step) {
        const ulonglong2 v = k2[i];
        count(v.x);
        count(v.y);
    }
    for (unsigned long long i = 2 * pairs + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) count(keys[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < ps.n_passes * RADIX; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// exclusive scan of every pass's histogram in place: hist[p][d] -> first output position of digit d in pass p
__global__ void __launch_bounds__(MAX_RADIX) radix_bases_kernel(uint32_t* __restrict__ hist) {
    constexpr int RADIX = MAX_RADIX;
    __shared__ uint32_t s[RADIX];
    uint32_t* h = hist + blockIdx.x * RADIX;
    const uint32_t v = h[threadIdx.x];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < RADIX; o <<= 1) {
        const uint32_t t = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0u;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    h[threadIdx.x] = s[threadIdx.x] - v;
}

// Lanes of the warp that hold the same digit.  MATCH.ANY does this in one instruction, but its throughput on B200 is one
// warp-instruction per ~58 cycles and SM for 8-bit digits (benchmarks/match_probe.cu: 0.65 ms per 100 M keys over the
// whole chip) -- exactly the time of a one-sweep pass, which made the sort MATCH-bound.  One ballot per digit bit and an
// AND of the matching halves gives the same mask 2.2x faster (0.30 ms per 100 M keys) on the regular pipes.
template <int RB>
__device__ __forceinline__ uint32_t digit_peers(uint32_t d) {
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < RB; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t v = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? v : ~v;
    }
    return peers;
}

template <bool PAIRS, int RB>
__global__ void __launch_bounds__(THREADS, 2)
radix_onesweep_kernel(const unsigned long long* __restrict__ keys_in, unsigned long long* __restrict__ keys_out,
                      const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out, uint32_t n, int shift, uint32_t mask,
                      const uint32_t* __restrict__ bin_base /* [RADIX] of this pass */, uint32_t* __restrict__ status /* [tiles][RADIX] */,
                      uint32_t* __restrict__ ticket) {
    constexpr int RADIX = 1 << RB;
    static_assert(RADIX <= THREADS, "one look-back thread per digit");
    extern __shared__ __align__(16) uint8_t rs_smem[];
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(rs_smem);                  // [TILE]
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(rs_smem + (size_t)TILE * 8);                    // [TILE] (pairs only)
    uint16_t* s_cnt = reinterpret_cast<uint16_t*>(rs_smem + (size_t)TILE * (PAIRS ? 12 : 8));      // [WARPS][RADIX], <= 8192
    uint32_t* s_base = reinterpret_cast<uint32_t*>(s_cnt + WARPS * RADIX);  // [RADIX] first slot of digit d in the sorted tile
    uint32_t* s_goff = s_base + RADIX;          // [RADIX] global position of sorted-tile slot 0 of digit d, minus s_base[d]
    uint32_t* s_scan = s_goff + RADIX;          // [WARPS] + tile id
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) s_scan[WARPS] = atomicAdd(ticket, 1u);
    for (uint32_t i = tid; i < WARPS * RADIX / 2; i += THREADS) reinterpret_cast<uint32_t*>(s_cnt)[i] = 0;
    __syncthreads();
    const uint32_t tile = s_scan[WARPS];
    const uint32_t base = tile * (uint32_t)TILE;
    const uint32_t n_tile = n - base < (uint32_t)TILE ? n - base : (uint32_t)TILE;

    // warp-striped load: warp w owns tile slots [w * 32 * IPT, (w + 1) * 32 * IPT), lane l its slots i * 32 + l
    unsigned long long key[IPT];
    uint16_t rank[IPT];
    const uint32_t wbase = warp * 32u * IPT;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const uint32_t slot = wbase + (uint32_t)i * 32u + lane;
        key[i] = slot < n_tile ? keys_in[base + slot] : ~0ull;  // padding: highest digit, behind every real key of the tile
    }
    // stable ranking inside the warp's chunk: rank = (same-digit keys of earlier rows of this warp) + (same-digit lanes
    // below me in this row).  The row's leader (lowest peer lane) bumps the warp-private counter -- two 16-bit counters
    // share a 32-bit word so that a plain 32-bit shared atomic serves both -- and broadcasts the old value.
    // (Plain load / leader store / __syncwarp instead of the atomic + shuffle was tried: it cost registers -- 100-240 bytes
    // of spills at the 64-register budget of two CTAs per SM -- and was 3 % (keys) to 14 % (pairs) slower.)
    uint16_t* my_cnt = s_cnt + warp * RADIX;
    uint32_t* my_cnt32 = reinterpret_cast<uint32_t*>(my_cnt);
    const uint32_t lt_mask = (1u << lane) - 1u;
    // GROUP rows are matched back to back (independent), then their counters are bumped in row order (shared atomics of
    // one warp on one address execute in program order, which keeps the ranking stable)
    constexpr int GROUP = 4;
#pragma unroll
    for (int i0 = 0; i0 < IPT; i0 += GROUP) {
        uint32_t d[GROUP], peers[GROUP];
#pragma unroll
        for (int g = 0; g < GROUP; ++g) {
            d[g] = (uint32_t)((key[i0 + g] >> shift) & mask);
            peers[g] = digit_peers<RB>(d[g]);
        }
        uint32_t old[GROUP];
#pragma unroll
        for (int g = 0; g < GROUP; ++g) {
            old[g] = 0;
            if ((peers[g] & lt_mask) == 0) old[g] = atomicAdd(&my_cnt32[d[g] >> 1], (uint32_t)__popc(peers[g]) << ((d[g] & 1u) * 16u));
        }
#pragma unroll
        for (int g = 0; g < GROUP; ++g) {
            const uint32_t o = __shfl_sync(0xffffffffu, old[g], __ffs((int)peers[g]) - 1);
            rank[i0 + g] = (uint16_t)(((o >> ((d[g] & 1u) * 16u)) & 0xFFFFu) + __popc(peers[g] & lt_mask));
        }
    }
    __syncthreads();
    // per digit: exclusive scan over the warps (in place), tile total
    uint32_t total = 0;
    if (tid < RADIX) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { const uint32_t c = s_cnt[w * RADIX + tid]; s_cnt[w * RADIX + tid] = (uint16_t)total; total += c; }
    }
    // padding keys all carry digit `mask` (all ones): they are not part of the tile's real counts
    const uint32_t n_pad = (uint32_t)TILE - n_tile;
    uint32_t real_total = total;
    if (tid == mask) real_total -= n_pad;
    // exclusive scan of the (padded) totals over the 256 digits -> position of each digit inside the sorted tile
    uint32_t incl = total;
    if (tid < RADIX) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
        if (lane == 31) s_scan[warp] = incl;
    }
    __syncthreads();
    uint32_t tile_base = 0;
    uint32_t* my_status = status + (size_t)tile * RADIX + tid;
    if (tid < RADIX) {
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; ++w) before += s_scan[w];
        tile_base = before + incl - total;
        s_base[tid] = tile_base;
        // publish the tile's own count at once: successors can already add it while this tile is still looking back
        *reinterpret_cast<volatile uint32_t*>(my_status) = (tile == 0 ? FLAG_INCLUSIVE : FLAG_LOCAL) | real_total;
    }
    __syncthreads();
    // reorder the tile in shared memory (needs only tile-local information) ...
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const uint32_t d = (uint32_t)((key[i] >> shift) & mask);
        const uint32_t pos = s_base[d] + my_cnt[d] + rank[i];
        s_keys[pos] = key[i];
        if (PAIRS) {  // the payload goes straight from global to its sorted slot (it never occupies registers)
            const uint32_t slot = wbase + (uint32_t)i * 32u + lane;
            s_vals[pos] = slot < n_tile ? vals_in[base + slot] : 0u;
        }
    }
    // ... then resolve the global offsets by decoupled look-back (one thread per digit)
    if (tid < RADIX) {
        uint32_t excl = 0;
        if (tile != 0) {
            // Speculative batches: the LOOK predecessors' words are loaded at once (independent L2 round trips instead of
            // a chain of them), then consumed in order; an unpublished word ends the batch and is polled again.
            int look = (int)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t s[LOOK];
#pragma unroll
                for (int b = 0; b < LOOK; ++b)
                    s[b] = look - b >= 0 ? *reinterpret_cast<const volatile uint32_t*>(status + (size_t)(look - b) * RADIX + tid) : FLAG_INCLUSIVE;
                int consumed = 0;
#pragma unroll
                for (int b = 0; b < LOOK; ++b) {
                    if (done || consumed != b) continue;
                    if (s[b] & FLAG_INCLUSIVE) { excl += s[b] & VALUE_MASK; done = true; }
                    else if (s[b] & FLAG_LOCAL) { excl += s[b] & VALUE_MASK; ++consumed; }
                }
                look -= consumed;
            }
            *reinterpret_cast<volatile uint32_t*>(my_status) = FLAG_INCLUSIVE | (excl + real_total);
        }
        s_goff[tid] = bin_base[tid] + excl - tile_base;
    }
    __syncthreads();
    // every digit's run leaves as one contiguous burst (padding occupies the last slots of the sorted tile)
    if (n_tile == (uint32_t)TILE) {  // full tile: four independent (key load -> offset load -> store) chains in flight
#pragma unroll
        for (int i0 = 0; i0 < IPT; i0 += 4) {
            unsigned long long k[4];
            uint32_t dst[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) k[g] = s_keys[tid + (uint32_t)(i0 + g) * THREADS];
#pragma unroll
            for (int g = 0; g < 4; ++g) dst[g] = s_goff[(uint32_t)((k[g] >> shift) & mask)] + tid + (uint32_t)(i0 + g) * THREADS;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                keys_out[dst[g]] = k[g];
                if (PAIRS) vals_out[dst[g]] = s_vals[tid + (uint32_t)(i0 + g) * THREADS];
            }
        }
    } else {
        for (uint32_t j = tid; j < n_tile; j += THREADS) {
            const unsigned long long k = s_keys[j];
            const uint32_t d = (uint32_t)((k >> shift) & mask);
            const uint32_t dst = s_goff[d] + j;
            keys_out[dst] = k;
            if (PAIRS) vals_out[dst] = s_vals[j];
        }
    }
}


}  // namespace rs

// exclusive scan of n counts in place by ONE CTA (tile counts: a 100 M-point cloud has 48 829 voxel-boundary tiles and
// 24 415 filter tiles), total -> *total_out
__global__ void __launch_bounds__(1024) scan_u32_kernel(uint32_t* __restrict__ counts, uint32_t n_tiles, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? counts[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((int)(threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sum[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if ((int)threadIdx.x >= o) w += y; }
            warp_sum[threadIdx.x] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t before = (threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0u;
        if (i < n_tiles) counts[i] = carry + before + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}

int exclusive_scan_u32(pb200_ctx* ctx, uint32_t* counts, uint32_t n, uint32_t* total_out) {
    scan_u32_kernel<<<1, 1024, 0, ctx->stream>>>(counts, n, total_out);
    g_launches++;
    PB_CUDA(cudaGetLastError());
    return PB200_OK;
}

// digit plan for the bits [begin_bit, end_bit): 9-bit digits only when they save a pass
static void plan_passes(const pb200_ctx* ctx, int begin_bit, int end_bit, bool pairs, rs::Passes* out, int* width) {
    using namespace rs;
    Passes& ps = *out;
    const int bits = end_bit - begin_bit;
    const int p8 = (bits + 7) / 8, p9 = (bits + 8) / 9;
    // (key, payload) tiles with 512 bins need 119 KB of shared memory: one CTA per SM instead of two, and a pass takes
    // 1.38 ms instead of 0.87 ms per 100 M pairs (measured) -- 9-bit digits are for keys-only sorts
    const bool wide = p9 < p8 && !pairs && !ctx->sort_force_8bit;
    ps.n_passes = wide ? p9 : p8;
    {   // spread the bits: the first passes are 9 bits wide as long as the rest still fills 8-bit passes, the last may be narrower
        int left = bits;
        for (int p = 0; p < ps.n_passes; ++p) {
            int w = wide ? ((left - 8 * (ps.n_passes - p - 1)) >= 9 ? 9 : 8) : 8;
            if (w > left) w = left;
            width[p] = w;
            left -= w;
        }
    }
    for (int p = 0; p < MAX_PASSES; ++p) { ps.shift[p] = 0; ps.mask[p] = 0; }
    for (int p = 0, sh = begin_bit; p < ps.n_passes; ++p) {
        ps.shift[p] = sh;
        ps.mask[p] = (1u << width[p]) - 1u;
        sh += width[p];
    }
}

// digit histograms of all passes of a sort over [begin_bit, end_bit), scanned into bin bases; layout [MAX_PASSES][MAX_RADIX]
// followed by one ticket counter per pass.
static int compute_hist(pb200_ctx* ctx, const unsigned long long* keys, uint64_t n, const rs::Passes& ps, DevTmp* d_hist) {
    using namespace rs;
    cudaStream_t st = ctx->stream;
    const size_t hist_bytes = (size_t)MAX_PASSES * MAX_RADIX * 4 + (MAX_PASSES + 4) * 4;
    PB_CUDA(d_hist->alloc(st, hist_bytes));
    PB_CUDA(cudaMemsetAsync(d_hist->p, 0, hist_bytes, st));
    unsigned long long hb = (n + 511) / 512, cap = (unsigned long long)ctx->sm_count * 4;
    PB_PHASE(ctx, "sort.histogram");
    radix_histogram_kernel<<<(unsigned)(hb < cap ? hb : cap), 512, 0, st>>>(keys, n, ps, (uint32_t*)d_hist->p);
    radix_bases_kernel<<<ps.n_passes, MAX_RADIX, 0, st>>>((uint32_t*)d_hist->p);
    g_launches += 2;
    PB_CUDA(cudaGetLastError());
    return PB200_OK;
}

// Sorts n keys by the bits [begin_bit, end_bit).  `keys` / `vals` are clobbered; the result ends up in (keys, vals) or in
// (keys_alt, vals_alt): *in_alt tells which.  vals / vals_alt may be null (keys only).
int radix_sort_u64(pb200_ctx* ctx, unsigned long long* keys, unsigned long long* keys_alt, uint32_t* vals, uint32_t* vals_alt,
                   uint64_t n, int begin_bit, int end_bit, bool* in_alt) {
    using namespace rs;
    *in_alt = false;
    if (n <= 1 || end_bit <= begin_bit) return PB200_OK;
    if (n >= (1ull << 30)) return set_error(PB200_ERR_UNSUPPORTED, "radix sort: more than 2^30 - 1 keys per call");
    if (end_bit - begin_bit > 64 || begin_bit < 0 || end_bit > 64) return set_error(PB200_ERR_INVALID, "radix sort: bad bit range");
    cudaStream_t st = ctx->stream;
    Passes ps;
    int width[MAX_PASSES];
    plan_passes(ctx, begin_bit, end_bit, vals != nullptr, &ps, width);
    const uint32_t n_tiles = (uint32_t)((n + TILE - 1) / TILE);
    DevTmp d_hist_own, d_status;
    DevTmp* d_hist = &d_hist_own;
    PB_TRY(compute_hist(ctx, keys, n, ps, d_hist));
    const size_t status_bytes = (size_t)ps.n_passes * n_tiles * MAX_RADIX * 4;
    PB_CUDA(d_status.alloc(st, status_bytes));
    PB_CUDA(cudaMemsetAsync(d_status.p, 0, status_bytes, st));
    uint32_t* hist = (uint32_t*)d_hist->p;
    uint32_t* tickets = hist + MAX_PASSES * MAX_RADIX;
    const bool pairs = vals != nullptr;
    auto smem_for = [&](int rb) { return (size_t)TILE * (pairs ? 12 : 8) + (size_t)WARPS * (1 << rb) * 2 + (size_t)(2 * (1 << rb) + WARPS + 4) * 4; };
    bool* attr_set = ctx->sort_attr_set;
    if (!attr_set[pairs ? 1 : 0]) {
        if (pairs) {
            PB_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(8)));
            PB_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<true, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(9)));
        } else {
            PB_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(8)));
            PB_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<false, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(9)));
        }
        attr_set[pairs ? 1 : 0] = true;
    }
    unsigned long long *src = keys, *dst = keys_alt;
    uint32_t *vsrc = vals, *vdst = vals_alt;
    for (int p = 0; p < ps.n_passes; ++p) {
        uint32_t* status = (uint32_t*)d_status.p + (size_t)p * n_tiles * MAX_RADIX;
        PB_PHASE(ctx, "sort.pass");
        const int rb = width[p] > 8 ? 9 : 8;
        const size_t smem = smem_for(rb);
#define PB_SWEEP(P, RB) radix_onesweep_kernel<P, RB><<<n_tiles, THREADS, smem, st>>>(src, dst, vsrc, vdst, (uint32_t)n, ps.shift[p], ps.mask[p], \
                                                                                  hist + p * MAX_RADIX, status, tickets + p)
        if (pairs) { if (rb == 9) PB_SWEEP(true, 9); else PB_SWEEP(true, 8); }
        else { if (rb == 9) PB_SWEEP(false, 9); else PB_SWEEP(false, 8); }
#undef PB_SWEEP
        g_launches++;
        unsigned long long* t = src; src = dst; dst = t;
        uint32_t* tv = vsrc; vsrc = vdst; vdst = tv;
    }
    PB_CUDA(cudaGetLastError());
    *in_alt = (ps.n_passes & 1) != 0;
    return PB200_OK;
}


}  // namespace pb200

using namespace pb200;

extern "C" int pb200_radix_sort_u64(pb200_ctx* ctx, uint64_t* keys, uint32_t* vals, uint64_t n, int begin_bit, int end_bit) {
    if (!ctx || (!keys && n)) return set_error(PB200_ERR_INVALID, "null argument");
    PB_DEVICE(ctx);
    if (n <= 1) return PB200_OK;
    DevTmp k2, v2;
    PB_CUDA(k2.alloc(ctx->stream, n * 8));
    if (vals) PB_CUDA(v2.alloc(ctx->stream, n * 4));
    bool in_alt = false;
    PB_TRY(radix_sort_u64(ctx, (unsigned long long*)keys, (unsigned long long*)k2.p, vals, (uint32_t*)v2.p, n, begin_bit, end_bit, &in_alt));
    if (in_alt) {
        PB_CUDA(cudaMemcpyAsync(keys, k2.p, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (vals) PB_CUDA(cudaMemcpyAsync(vals, v2.p, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return PB200_OK;
}
