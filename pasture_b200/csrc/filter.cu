// filter.cu -- HashMapBuffer::filter / filter_into (pasture-core/src/containers/point_buffer.rs:1064-1136) as a
// stream compaction in three launches: per-tile match counts (4096 points per tile, 1 B/point), an exclusive scan of
// the tile counts, and one compaction kernel in which every CTA rebuilds the in-tile ranks from the mask, keeps the
// list of surviving local indices in shared memory and then copies every byte stream (whole records, or one stream per
// attribute) in OUTPUT order: consecutive threads write consecutive words of the target, so the writes are fully
// coalesced for any record stride and the gathered reads stay inside the tile.
// Algorithmic bytes per point: 2 (mask, read twice) + stride (source) + kept fraction x stride (target).  The reference takes a closure `Fn(usize) -> bool` over the point index; the
// FFI-crossable form of that is a byte mask with one entry per point (non-zero = keep).
#include "internal.h"

namespace pb200 {

constexpr int FT_THREADS = 256;
constexpr int FT_PPT = 16;                       // consecutive points per thread
constexpr int FT_TILE = FT_THREADS * FT_PPT;     // 4096 points per tile
constexpr int FT_MAX_STREAMS = PB200_MAX_ATTRIBUTES;

// one byte stream of the compaction: `wps` words of W bytes per point on both sides
struct FStream {
    const uint8_t* src;
    uint8_t* dst;
    uint32_t sstride, dstride;
    uint32_t wbytes, wps;
};
struct FPlan {
    FStream s[FT_MAX_STREAMS];
    int n_streams;
};

__device__ __forceinline__ uint32_t load_mask16(const uint8_t* __restrict__ mask, unsigned long long i0, unsigned long long n, bool aligned) {
    // -> bit j set iff mask[i0 + j] != 0, j < 16
    uint32_t bits = 0;
    if (aligned && i0 + 16 <= n) {
        const uint4 v = *reinterpret_cast<const uint4*>(mask + i0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t ne = __vcmpne4(w[k], 0u);  // 0xFF per non-zero byte
            bits |= (((ne >> 7) & 1u) | ((ne >> 14) & 2u) | ((ne >> 21) & 4u) | ((ne >> 28) & 8u)) << (4 * k);
        }
    } else {
        for (int j = 0; j < 16; ++j)
            if (i0 + j < n && mask[i0 + j]) bits |= 1u << j;
    }
    return bits;
}

__global__ void __launch_bounds__(FT_THREADS) filter_count_kernel(const uint8_t* __restrict__ mask, unsigned long long n,
                                                                  uint32_t n_tiles, uint32_t* __restrict__ tile_counts) {
    __shared__ uint32_t s_warp[FT_THREADS / 32];
    const bool aligned = ((uintptr_t)mask & 15) == 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long i0 = (unsigned long long)tile * FT_TILE + (unsigned long long)threadIdx.x * FT_PPT;
        uint32_t c = i0 < n ? __popc(load_mask16(mask, i0, n, aligned)) : 0;
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < FT_THREADS / 32; ++w) t += s_warp[w];
            tile_counts[tile] = t;
        }
        __syncthreads();
    }
}

template <class W>
__device__ __forceinline__ void compact_stream(const FStream& st, unsigned long long tile0, unsigned long long out0, uint32_t cnt,
                                               const uint16_t* __restrict__ kept) {
    const uint32_t wps = st.wps, total = cnt * wps;
    uint32_t p = threadIdx.x / wps, w = threadIdx.x - p * wps;
    const uint32_t dp = FT_THREADS / wps, dw = FT_THREADS - dp * wps;
    for (uint32_t j = threadIdx.x; j < total; j += FT_THREADS) {
        const W v = reinterpret_cast<const W*>(st.src + (tile0 + kept[p]) * st.sstride)[w];
        reinterpret_cast<W*>(st.dst + (out0 + p) * st.dstride)[w] = v;
        p += dp; w += dw;
        if (w >= wps) { w -= wps; p++; }
    }
}

// tile_offsets = exclusive scan of tile_counts
__global__ void __launch_bounds__(FT_THREADS) filter_compact_kernel(const uint8_t* __restrict__ mask, unsigned long long n, uint32_t n_tiles,
                                                                    const uint32_t* __restrict__ tile_offsets,
                                                                    const __grid_constant__ FPlan plan) {
    __shared__ uint16_t s_kept[FT_TILE];
    __shared__ uint32_t s_warp[FT_THREADS / 32];
    const bool aligned = ((uintptr_t)mask & 15) == 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long tile0 = (unsigned long long)tile * FT_TILE;
        const unsigned long long i0 = tile0 + (unsigned long long)threadIdx.x * FT_PPT;
        const uint32_t bits = i0 < n ? load_mask16(mask, i0, n, aligned) : 0;
        const uint32_t c = __popc(bits);
        uint32_t incl = c;  // warp inclusive scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t base = 0, cnt = 0;
#pragma unroll
        for (int w = 0; w < FT_THREADS / 32; ++w) {
            const uint32_t t = s_warp[w];
            if (w < warp) base += t;
            cnt += t;
        }
        uint32_t r = base + incl - c;
        uint32_t b = bits;
        while (b) {
            const int j = __ffs(b) - 1;
            b &= b - 1;
            s_kept[r++] = (uint16_t)(threadIdx.x * FT_PPT + j);
        }
        __syncthreads();
        const unsigned long long out0 = tile_offsets[tile];
        for (int si = 0; si < plan.n_streams; ++si) {
            const FStream& st = plan.s[si];
            switch (st.wbytes) {
                case 16: compact_stream<uint4>(st, tile0, out0, cnt, s_kept); break;
                case 8: compact_stream<unsigned long long>(st, tile0, out0, cnt, s_kept); break;
                case 4: compact_stream<uint32_t>(st, tile0, out0, cnt, s_kept); break;
                case 2: compact_stream<uint16_t>(st, tile0, out0, cnt, s_kept); break;
                default: compact_stream<uint8_t>(st, tile0, out0, cnt, s_kept); break;
            }
        }
        __syncthreads();  // s_kept / s_warp are reused by the next tile
    }
}

using FBuf = DevTmp;  // stream-ordered temporaries, recycled between calls

static uint32_t word_bytes(uint64_t size, const void* sp, uint64_t ss, const void* dp, uint64_t ds) {
    uint32_t w = 16;
    while (w > 1 && ((size % w) || ((uintptr_t)sp % w) || (ss % w) || ((uintptr_t)dp % w) || (ds % w))) w >>= 1;
    return w;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_filter_into(pb200_ctx* ctx, const pb200_buffer_desc* src, const uint8_t* mask, const pb200_buffer_desc* dst,
                      uint64_t* num_matches) {
    if (!ctx || !num_matches) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(src, "source buffer"));
    PB_TRY(validate_desc(dst, "target buffer"));
    if (!pb200_layout_equal(src->layout, dst->layout)) return set_error(PB200_ERR_LAYOUT_MISMATCH, "PointLayouts must match");  // :1092-1094
    *num_matches = 0;
    const uint64_t n = src->len;
    if (n == 0) return PB200_OK;
    if (!mask) return set_error(PB200_ERR_INVALID, "null mask");
    if (n > 0xFFFFFFFFull) return set_error(PB200_ERR_UNSUPPORTED, "more than 2^32-1 points per call");
    PB_DEVICE(ctx);
    cudaStream_t st = ctx->stream;
    const pb200_layout& L = *src->layout;
    // device views of the source (the mask lives in the source's memory space)
    FBuf d_mask_buf, d_src_aos;
    std::vector<FBuf> d_src_cols(L.attrs.size());
    const uint8_t* d_mask = mask;
    const uint8_t* s_aos = (const uint8_t*)src->aos;
    std::vector<const uint8_t*> s_cols(L.attrs.size(), nullptr);
    if (src->kind == PB200_COLUMNAR)
        for (size_t a = 0; a < L.attrs.size(); ++a) s_cols[a] = (const uint8_t*)src->columns[a];
    if (src->memspace == PB200_HOST) {
        PB_CUDA(d_mask_buf.alloc(st, (size_t)n));
        PB_CUDA(cudaMemcpyAsync(d_mask_buf.p, mask, (size_t)n, cudaMemcpyHostToDevice, st));
        d_mask = (const uint8_t*)d_mask_buf.p;
        if (src->kind == PB200_INTERLEAVED) {
            PB_CUDA(d_src_aos.alloc(st, (size_t)(n * L.size)));
            PB_CUDA(cudaMemcpyAsync(d_src_aos.p, src->aos, (size_t)(n * L.size), cudaMemcpyHostToDevice, st));
            s_aos = (const uint8_t*)d_src_aos.p;
        } else {
            for (size_t a = 0; a < L.attrs.size(); ++a) {
                if (L.attrs[a].size == 0) continue;
                PB_CUDA(d_src_cols[a].alloc(st, (size_t)(n * L.attrs[a].size)));
                PB_CUDA(cudaMemcpyAsync(d_src_cols[a].p, src->columns[a], (size_t)(n * L.attrs[a].size), cudaMemcpyHostToDevice, st));
                s_cols[a] = (const uint8_t*)d_src_cols[a].p;
            }
        }
    }
    // per-tile match counts -> exclusive scan -> total
    const uint32_t n_tiles = (uint32_t)((n + FT_TILE - 1) / FT_TILE);
    FBuf d_offsets;  // per-tile counts, scanned in place into the tiles' output offsets; one extra word for the total
    PB_CUDA(d_offsets.alloc(st, ((size_t)n_tiles + 1) * 4));
    const uint32_t cap = (uint32_t)ctx->sm_count * 8;
    const uint32_t blocks = n_tiles < cap ? n_tiles : cap;
    filter_count_kernel<<<blocks, FT_THREADS, 0, st>>>(d_mask, n, n_tiles, (uint32_t*)d_offsets.p);
    g_launches++;
    PB_TRY(exclusive_scan_u32(ctx, (uint32_t*)d_offsets.p, n_tiles, (uint32_t*)d_offsets.p + n_tiles));
    uint32_t last[2] = {0, 0};
    PB_CUDA(cudaMemcpyAsync(&last[0], (uint32_t*)d_offsets.p + n_tiles, 4, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    const uint64_t m = (uint64_t)last[0] + last[1];
    *num_matches = m;
    if (dst->len < m)  // :1097-1099
        return set_error(PB200_ERR_RANGE, "buffer.len() must be at least as large as the number of predicate matches");
    if (m == 0) return PB200_OK;
    // device views of the target
    FBuf d_dst_aos;
    std::vector<FBuf> d_dst_cols(L.attrs.size());
    uint8_t* t_aos = (uint8_t*)dst->aos;
    std::vector<uint8_t*> t_cols(L.attrs.size(), nullptr);
    if (dst->kind == PB200_COLUMNAR)
        for (size_t a = 0; a < L.attrs.size(); ++a) t_cols[a] = (uint8_t*)dst->columns[a];
    if (dst->memspace == PB200_HOST) {
        if (dst->kind == PB200_INTERLEAVED) {  // padding / bytes between attributes must survive: start from the target's bytes
            PB_CUDA(d_dst_aos.alloc(st, (size_t)(m * L.size)));
            PB_CUDA(cudaMemcpyAsync(d_dst_aos.p, dst->aos, (size_t)(m * L.size), cudaMemcpyHostToDevice, st));
            t_aos = (uint8_t*)d_dst_aos.p;
        } else {
            for (size_t a = 0; a < L.attrs.size(); ++a) {
                if (L.attrs[a].size == 0) continue;
                PB_CUDA(d_dst_cols[a].alloc(st, (size_t)(m * L.attrs[a].size)));
                t_cols[a] = (uint8_t*)d_dst_cols[a].p;
            }
        }
    }
    // byte streams: whole records when both sides are interleaved and the layout has no padding, else one per attribute
    FPlan plan{};
    uint64_t attr_bytes = 0;
    bool contiguous = true;
    for (const pb200_attr& at : L.attrs) {
        if (at.offset != attr_bytes) contiguous = false;
        attr_bytes += at.size;
    }
    auto add_stream = [&](const uint8_t* sp, uint64_t ss, uint8_t* dp, uint64_t ds, uint64_t size) {
        FStream& f = plan.s[plan.n_streams++];
        f.src = sp; f.dst = dp; f.sstride = (uint32_t)ss; f.dstride = (uint32_t)ds;
        f.wbytes = word_bytes(size, sp, ss, dp, ds);
        f.wps = (uint32_t)(size / f.wbytes);
    };
    if (src->kind == PB200_INTERLEAVED && dst->kind == PB200_INTERLEAVED && contiguous && attr_bytes == L.size && L.size) {
        add_stream(s_aos, L.size, t_aos, L.size, L.size);
    } else {
        for (size_t a = 0; a < L.attrs.size(); ++a) {
            const pb200_attr& at = L.attrs[a];
            if (at.size == 0) continue;
            if (at.size > 0xFFFFFFFFull || L.size > 0xFFFFFFFFull) return set_error(PB200_ERR_UNSUPPORTED, "attribute too large");
            add_stream(src->kind == PB200_INTERLEAVED ? s_aos + at.offset : s_cols[a], src->kind == PB200_INTERLEAVED ? L.size : at.size,
                       dst->kind == PB200_INTERLEAVED ? t_aos + at.offset : t_cols[a], dst->kind == PB200_INTERLEAVED ? L.size : at.size,
                       at.size);
        }
    }
    if (plan.n_streams) {
        filter_compact_kernel<<<blocks, FT_THREADS, 0, st>>>(d_mask, n, n_tiles, (const uint32_t*)d_offsets.p, plan);
        g_launches++;
    }
    PB_CUDA(cudaGetLastError());
    if (dst->memspace == PB200_HOST) {
        if (dst->kind == PB200_INTERLEAVED) PB_CUDA(cudaMemcpyAsync(dst->aos, t_aos, (size_t)(m * L.size), cudaMemcpyDeviceToHost, st));
        else
            for (size_t a = 0; a < L.attrs.size(); ++a)
                if (L.attrs[a].size) PB_CUDA(cudaMemcpyAsync(dst->columns[a], t_cols[a], (size_t)(m * L.attrs[a].size), cudaMemcpyDeviceToHost, st));
    }
    PB_CUDA(cudaStreamSynchronize(st));  // temporaries die here
    return PB200_OK;
}

}  // extern "C"
