// synth.cu -- deterministic synthetic inputs generated directly in HBM: the same splitmix64
// streams the CPU checker generates, so bench and parity tests can build
// 100 M-point clouds without a host round trip.  Bench/test tooling, not part of the conversion path.
#include "internal.h"

namespace pb200 {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long seed, unsigned long long j) {
    unsigned long long z = seed + (j + 1) * 0x9E3779B97F4A7C15ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
__device__ __forceinline__ double u01(unsigned long long h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

// raw LAS format-0 records (20 B): one thread per point, 5 x 32-bit stores (records are 4 B aligned)
__global__ void __launch_bounds__(256) synth_las_fmt0_kernel(uint32_t* __restrict__ out, unsigned long long first,
                                                             unsigned long long n, unsigned long long seed) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += step) {
        const unsigned long long i = first + k;
        uint32_t w[5];
        for (int c = 0; c < 3; ++c) w[c] = (uint32_t)((int)(splitmix64(seed, 8 * i + c) % 2000001ull) - 1000000);
        const unsigned long long h3 = splitmix64(seed, 8 * i + 3), h4 = splitmix64(seed, 8 * i + 4), h5 = splitmix64(seed, 8 * i + 5);
        w[3] = (uint32_t)(h3 >> 48) | ((uint32_t)(h4 & 0xFF) << 16) | ((uint32_t)((h4 >> 8) & 0xFF) << 24);
        w[4] = (uint32_t)((h4 >> 16) & 0xFF) | ((uint32_t)((h4 >> 24) & 0xFF) << 8) | ((uint32_t)(h5 & 0xFFFF) << 16);
        uint32_t* r = out + 5 * k;
        for (int j = 0; j < 5; ++j) r[j] = w[j];
    }
}

__global__ void __launch_bounds__(256) synth_terrain_kernel(double* __restrict__ out, unsigned long long first,
                                                            unsigned long long n, unsigned long long seed) {
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += step) {
        const unsigned long long i = first + k;
        const double x = __dmul_rn(u01(splitmix64(seed, 8 * i + 0)), 500.0);
        const double y = __dmul_rn(u01(splitmix64(seed, 8 * i + 1)), 500.0);
        const double a = __dmul_rn(x, 0.002), b = __dmul_rn(y, 0.002);
        const double t1 = __dmul_rn(10.0, __dsub_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
        const double t2 = __dmul_rn(5.0, __dmul_rn(a, b));
        const double t3 = __dmul_rn(0.1, u01(splitmix64(seed, 8 * i + 2)));
        out[3 * k] = x;
        out[3 * k + 1] = y;
        out[3 * k + 2] = __dadd_rn(__dadd_rn(t1, t2), t3);
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_synth_las_fmt0_records(pb200_ctx* ctx, void* device_out, uint64_t first_index, uint64_t n, uint64_t seed) {
    PB_DEVICE(ctx);
    if (!device_out || ((uintptr_t)device_out & 3)) return set_error(PB200_ERR_INVALID, "output must be a 4-byte aligned device pointer");
    if (n == 0) return PB200_OK;
    unsigned long long want = (n + 255) / 256, cap = (unsigned long long)ctx->sm_count * 16;
    synth_las_fmt0_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>((uint32_t*)device_out, first_index, n, seed);
    PB_CUDA(cudaGetLastError());
    return PB200_OK;
}

int pb200_synth_terrain_positions(pb200_ctx* ctx, void* device_out, uint64_t first_index, uint64_t n, uint64_t seed) {
    PB_DEVICE(ctx);
    if (!device_out || ((uintptr_t)device_out & 7)) return set_error(PB200_ERR_INVALID, "output must be an 8-byte aligned device pointer");
    if (n == 0) return PB200_OK;
    unsigned long long want = (n + 255) / 256, cap = (unsigned long long)ctx->sm_count * 16;
    synth_terrain_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>((double*)device_out, first_index, n, seed);
    PB_CUDA(cudaGetLastError());
    return PB200_OK;
}

}  // extern "C"
