// filter.cu -- HashMapBuffer::filter / filter_into (pasture-core/src/containers/point_buffer.rs:1064-1136) as a
// stream compaction: exclusive scan of the predicate mask, then one gather/scatter pass per attribute into an
// interleaved or columnar target.  The reference takes a closure `Fn(usize) -> bool` over the point index; the
// FFI-crossable form of that is a byte mask with one entry per point (non-zero = keep).
#include <cub/device/device_scan.cuh>

#include "internal.h"

namespace pb200 {

__global__ void __launch_bounds__(256) mask_to_flags_kernel(const uint8_t* __restrict__ mask, unsigned long long n,
                                                            uint32_t* __restrict__ flags) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        flags[i] = mask[i] ? 1u : 0u;
}

// one thread per (kept point, W-sized word of the attribute)
template <class W>
__global__ void __launch_bounds__(256) compact_attribute_kernel(const uint8_t* __restrict__ mask, const uint32_t* __restrict__ pos,
                                                                unsigned long long n, const uint8_t* __restrict__ src,
                                                                unsigned long long sstride, uint8_t* __restrict__ dst,
                                                                unsigned long long dstride, uint32_t words) {
    const unsigned long long total = n * words;
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += step) {
        const unsigned long long i = t / words;
        const uint32_t w = (uint32_t)(t - i * words);
        if (!mask[i]) continue;
        reinterpret_cast<W*>(dst + (unsigned long long)pos[i] * dstride)[w] = reinterpret_cast<const W*>(src + i * sstride)[w];
    }
}

struct FBuf {
    void* p = nullptr;
    ~FBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t b) { return cudaMalloc(&p, b ? b : 1); }
};

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
    PB_TRY(ensure_device(ctx));
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
        PB_CUDA(d_mask_buf.alloc((size_t)n));
        PB_CUDA(cudaMemcpyAsync(d_mask_buf.p, mask, (size_t)n, cudaMemcpyHostToDevice, st));
        d_mask = (const uint8_t*)d_mask_buf.p;
        if (src->kind == PB200_INTERLEAVED) {
            PB_CUDA(d_src_aos.alloc((size_t)(n * L.size)));
            PB_CUDA(cudaMemcpyAsync(d_src_aos.p, src->aos, (size_t)(n * L.size), cudaMemcpyHostToDevice, st));
            s_aos = (const uint8_t*)d_src_aos.p;
        } else {
            for (size_t a = 0; a < L.attrs.size(); ++a) {
                if (L.attrs[a].size == 0) continue;
                PB_CUDA(d_src_cols[a].alloc((size_t)(n * L.attrs[a].size)));
                PB_CUDA(cudaMemcpyAsync(d_src_cols[a].p, src->columns[a], (size_t)(n * L.attrs[a].size), cudaMemcpyHostToDevice, st));
                s_cols[a] = (const uint8_t*)d_src_cols[a].p;
            }
        }
    }
    // scan
    FBuf d_flags, d_pos, d_tmp;
    PB_CUDA(d_flags.alloc((size_t)n * 4));
    PB_CUDA(d_pos.alloc((size_t)n * 4));
    const unsigned long long cap = (unsigned long long)ctx->sm_count * 16;
    const unsigned blocks = (unsigned)(((n + 255) / 256) < cap ? ((n + 255) / 256) : cap);
    mask_to_flags_kernel<<<blocks, 256, 0, st>>>(d_mask, n, (uint32_t*)d_flags.p);
    g_launches++;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const uint32_t*)d_flags.p, (uint32_t*)d_pos.p, (int)n, st);
    PB_CUDA(d_tmp.alloc(tmp_bytes));
    PB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, (const uint32_t*)d_flags.p, (uint32_t*)d_pos.p, (int)n, st));
    g_launches++;
    uint32_t last[2];
    PB_CUDA(cudaMemcpyAsync(&last[0], (uint32_t*)d_pos.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaMemcpyAsync(&last[1], (uint32_t*)d_flags.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
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
            PB_CUDA(d_dst_aos.alloc((size_t)(m * L.size)));
            PB_CUDA(cudaMemcpyAsync(d_dst_aos.p, dst->aos, (size_t)(m * L.size), cudaMemcpyHostToDevice, st));
            t_aos = (uint8_t*)d_dst_aos.p;
        } else {
            for (size_t a = 0; a < L.attrs.size(); ++a) {
                if (L.attrs[a].size == 0) continue;
                PB_CUDA(d_dst_cols[a].alloc((size_t)(m * L.attrs[a].size)));
                t_cols[a] = (uint8_t*)d_dst_cols[a].p;
            }
        }
    }
    for (size_t a = 0; a < L.attrs.size(); ++a) {
        const pb200_attr& at = L.attrs[a];
        if (at.size == 0) continue;
        const uint8_t* sp = src->kind == PB200_INTERLEAVED ? s_aos + at.offset : s_cols[a];
        const uint64_t ss = src->kind == PB200_INTERLEAVED ? L.size : at.size;
        uint8_t* dp = dst->kind == PB200_INTERLEAVED ? t_aos + at.offset : t_cols[a];
        const uint64_t ds = dst->kind == PB200_INTERLEAVED ? L.size : at.size;
        uint32_t w = 8;
        while (w > 1 && ((at.size % w) || ((uintptr_t)sp % w) || (ss % w) || ((uintptr_t)dp % w) || (ds % w))) w >>= 1;
        const uint32_t words = (uint32_t)(at.size / w);
        const unsigned long long total = n * words;
        const unsigned b2 = (unsigned)(((total + 255) / 256) < cap * 4 ? ((total + 255) / 256) : cap * 4);
        if (w == 8) compact_attribute_kernel<unsigned long long><<<b2, 256, 0, st>>>(d_mask, (const uint32_t*)d_pos.p, n, sp, ss, dp, ds, words);
        else if (w == 4) compact_attribute_kernel<uint32_t><<<b2, 256, 0, st>>>(d_mask, (const uint32_t*)d_pos.p, n, sp, ss, dp, ds, words);
        else if (w == 2) compact_attribute_kernel<uint16_t><<<b2, 256, 0, st>>>(d_mask, (const uint32_t*)d_pos.p, n, sp, ss, dp, ds, words);
        else compact_attribute_kernel<uint8_t><<<b2, 256, 0, st>>>(d_mask, (const uint32_t*)d_pos.p, n, sp, ss, dp, ds, words);
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
