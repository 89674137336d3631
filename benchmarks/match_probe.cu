// match_probe.cu -- what does a warp-level "which lanes hold my digit" cost on B200?  The one-sweep radix sort ranks
// its keys with MATCH.ANY + a leader atomic; ncu shows 42 % of the stall samples on the instruction that consumes the
// MATCH result.  This probe times, per key and SM, (a) match.any on 8-bit digits, (b) the same peers mask from eight
// ballots (bit peeling), (c) match.any on 4-bit digits (fewer distinct values), with full occupancy and 16 independent
// keys per thread, no memory traffic.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o benchmarks/build/match_probe benchmarks/match_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int MODE, int BITS>
__global__ void __launch_bounds__(512) probe(uint32_t* out, int iters) {
    uint32_t acc = 0, seed = blockIdx.x * 7919u + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        uint32_t d[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { seed = mix32(seed + i); d[i] = seed & ((1u << BITS) - 1u); }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint32_t peers;
            if (MODE == 0) {
                peers = __match_any_sync(0xffffffffu, d[i]);
            } else {
                peers = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < BITS; ++b) {
                    const uint32_t bit = (d[i] >> b) & 1u;
                    const uint32_t v = __ballot_sync(0xffffffffu, bit);
                    peers &= bit ? v : ~v;
                }
            }
            acc += __popc(peers & ((1u << lane) - 1u));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE, int BITS>
static void run(const char* name, uint32_t* out, int sms) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE, BITS><<<sms * 2, 512>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE, BITS><<<sms * 2, 512>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double keys_per_sm = 2.0 * 512 * 16 * iters;
    printf("{\"variant\": \"%s\", \"ms\": %.3f, \"ns_per_key_per_sm\": %.4f, \"ms_per_100M_keys_148sm\": %.4f}\n", name, ms,
           ms * 1e6 / keys_per_sm, ms * 1e6 / keys_per_sm * 1e8 / 148 * 1e-6);
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out;
    cudaMalloc(&out, (size_t)sms * 2 * 512 * 4);
    run<0, 8>("match.any 8-bit digits (+ hash + popc)", out, sms);
    run<1, 8>("8 ballots 8-bit digits (+ hash + popc)", out, sms);
    run<0, 4>("match.any 4-bit digits", out, sms);
    run<1, 4>("4 ballots 4-bit digits", out, sms);
    run<0, 9>("match.any 9-bit digits", out, sms);
    run<1, 9>("9 ballots 9-bit digits", out, sms);
    run<1, 1>("1 ballot (hash + popc floor)", out, sms);
    return 0;
}
